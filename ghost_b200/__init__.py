"""ghost_b200: B200-native Morse-wavelet continuous wavelet transform.

Drop-in for the ``ghost.wave.ContinuousWaveletTransform`` hot path of nelpy/ghost.
All arithmetic runs in hand-written CUDA (sm_100a) behind a C ABI
(``include/ghost_cwt.h``, built into ``ghost_b200/libghostcwt.so``); there is no CPU
fallback.
"""
from . import utils
from . import formats
from . import wave
from . import sigtools
from .wave import ContinuousWaveletTransform, Morse

__version__ = "0.1.0"
__all__ = ["ContinuousWaveletTransform", "Morse", "wave", "sigtools", "formats", "utils"]

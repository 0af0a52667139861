from .wavelet import Wavelet
from .morse import Morse, morsefreq, morsehigh
from .transforms import ContinuousWaveletTransform

__all__ = ["Wavelet", "Morse", "ContinuousWaveletTransform", "morsefreq", "morsehigh"]

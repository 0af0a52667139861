"""Device versions of ghost.sigtools (reference ghost/sigtools/__init__.py exports the same
names).  The ``*_scipy`` / ``*_fftw`` pairs of the reference differ only in the CPU FFT library
behind them; here both names run the same CUDA code."""
from .convolution import fastconv, fastconv_scipy, fastconv_fftw, fastconv_freq_scipy, fastconv_freq_fftw
from .fourier import chirpz_dft, dft
from .analytic import analytic_signal, analytic_signal_scipy, analytic_signal_fftw

__all__ = ["fastconv", "fastconv_scipy", "fastconv_fftw", "fastconv_freq_scipy", "fastconv_freq_fftw",
           "chirpz_dft", "dft", "analytic_signal", "analytic_signal_scipy", "analytic_signal_fftw"]

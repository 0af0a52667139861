"""Abstract wavelet type (reference ghost/wave/wavelet.py:7-21)."""
from abc import ABC, abstractmethod

__all__ = ["Wavelet"]


class Wavelet(ABC):

    def __init__(self):
        pass

    def __repr__(self):
        return self.__class__.__name__

    @abstractmethod
    def copy(self):
        """Independent copy of the wavelet object."""
        return

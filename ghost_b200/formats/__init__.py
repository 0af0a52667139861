from .preprocessing import standardize_input
from .postprocessing import output_numpy_or_asa

__all__ = ["standardize_input", "output_numpy_or_asa"]

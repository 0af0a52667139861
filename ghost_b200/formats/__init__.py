from .preprocessing import standardize_input

__all__ = ["standardize_input"]

"""custrings_b200 — B200-native drop-in for the hot path of rapidsai/custrings.

    from custrings_b200 import nvstrings, nvcategory, nvtext
"""
from . import nvstrings  # noqa: F401

__all__ = ["nvstrings", "nvcategory", "nvtext"]


def __getattr__(name):
    if name in ("nvcategory", "nvtext"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)

"""Shim of the two torch_geometric 1.6.3 symbols the reference hot path touches."""
from . import nn, data  # noqa: F401

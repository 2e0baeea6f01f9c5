"""Greedy sparse-NNLS solvers (reference: bayesiancoresets/snnls/__init__.py:1-4).  The
sampling baselines of the reference (snnls/sampling.py) are out of scope of this engine."""
from .base import SparseNNLS
from .giga import GIGA, FrankWolfe
from .orthopursuit import OrthoPursuit

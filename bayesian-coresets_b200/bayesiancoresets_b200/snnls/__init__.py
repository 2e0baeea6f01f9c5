"""Sparse-NNLS solvers (reference: bayesiancoresets/snnls/__init__.py:1-4): the greedy solvers run on the device; the
sampling baselines are O(1) per iteration and only use the device for error() / optimize()."""
from .base import SparseNNLS
from .giga import GIGA, FrankWolfe
from .orthopursuit import OrthoPursuit
from .sampling import ImportanceSampling, UniformSampling

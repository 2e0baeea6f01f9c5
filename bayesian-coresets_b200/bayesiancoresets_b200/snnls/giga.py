import numpy as np
from .base import SparseNNLS
from .. import _native as nat


class GIGA(SparseNNLS):
  """Greedy iterative geodesic ascent on the device (reference: snnls/giga.py:6-64).

  Per iteration ONE float32 pass over the unit rows computes <a_n, cdir> and <a_n, xw>, the
  geodesic score and its argmax (scan kernel); the closed-form line search and the reweight
  run in float64 on the K stored active rows (step kernel)."""
  _alg = nat.ALG_GIGA


class FrankWolfe(SparseNNLS):
  """Frank-Wolfe on the scaled simplex (reference: snnls/frankwolfe.py:5-40)."""
  _alg = nat.ALG_FW

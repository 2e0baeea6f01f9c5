from .base import SparseNNLS
from .. import _native as nat


class OrthoPursuit(SparseNNLS):
  """Orthogonal matching pursuit (reference: snnls/orthopursuit.py:7-42).

  Everything runs on the device, `build(itrs)` without a host round trip per iteration: the
  N x S residual-correlation scan, the negative direction over the active set, and the NNLS
  re-solve on the K active columns (float64 Lawson-Hanson, warm-started from the previous passive
  set -- csrc/nnls_logic.h; the NNLS minimiser is unique, so it equals `scipy.optimize.nnls`;
  tests/scipy_omp.py drives the same device primitives with the reference's SciPy call as a cross-check)."""
  _alg = nat.ALG_OMP

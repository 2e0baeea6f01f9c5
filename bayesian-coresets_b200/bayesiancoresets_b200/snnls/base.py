"""GPU-backed sparse-NNLS solver base: same surface as the reference's SparseNNLS
(bayesiancoresets/snnls/snnls.py:8-106) -- ctor (A, b), build / weights / error / size /
optimize / reset, attributes w / A / b / reached_numeric_limit -- with the greedy loop running
on the device through the C-ABI (bcg_solver_build)."""
import logging
import secrets
import struct
import numpy as np

from .. import util
from ..util import NumericalPrecisionError
from .. import _native as nat
from ..comm import SerialComm


class SparseNNLS(object):
  _alg = None

  def __init__(self, A, b, check_error_monotone=True, comm=None):
    self.alg_name = self.__class__.__name__ + '-' + secrets.token_hex(3)
    self.log = logging.LoggerAdapter(logging.getLogger(), {'id': self.alg_name})
    self.check_error_monotone = bool(check_error_monotone)
    self.comm = comm or SerialComm()
    if isinstance(A, nat.DeviceVecsT):
      vecs = A.vecs                                   # already on the device: vecs.T of a GPU projector
    else:
      A = np.asarray(A)
      if A.ndim != 2:
        raise ValueError('A must be (S, N)')
      vecs = nat.DeviceVecs.from_host(A.T)            # (N, S); zero-copy when A is the F-ordered vecs.T view
    self.A = A
    self.b = np.asarray(b, dtype=np.float64)
    self._vecs = vecs
    self.reached_numeric_limit = False
    # One collective for the whole bootstrap: (local rows, zero rows, sum of local norms, mailbox handle) of every
    # rank in a fixed-size blob.  Every rank derives the same layout / totals from it (sums in rank order), and a
    # zero column anywhere raises on EVERY rank -- no rank is left waiting in a collective.
    nloc, zloc, nsum_loc = vecs.shape[0], vecs.zero_rows(), vecs.norm_sum()
    if self.comm.world > 1:
      blob = struct.pack('<qqd', nloc, zloc, nsum_loc) + vecs.ctx.comm_handle()
      parts = self.comm.allgather_bytes(blob)
      meta = [struct.unpack('<qqd', p[:24]) for p in parts]
      handles = [p[24:88] for p in parts]
    else:
      meta, handles = [(nloc, zloc, nsum_loc)], None
    self._counts = [int(m[0]) for m in meta]
    self.row_offset, self.n_global = sum(self._counts[:self.comm.rank]), sum(self._counts)
    self._check_zero_columns(sum(int(m[1]) for m in meta))
    norm_sum = 0.
    for m in meta:
      norm_sum += m[2]
    try:
      self._native = nat.NativeSolver(vecs, self._alg, self.b, norm_sum, self.row_offset, self.n_global)
    except nat.BcgError as e:
      if e.code == nat.ERR_ZERO_B:
        raise NumericalPrecisionError('norm of b must be > 0')
      raise
    if not self.check_error_monotone:
      self._native.set_check_monotone(False)          # snnls.py:46,56: no monotone test, retry flag never cleared
    if self.comm.world > 1:
      self._native.comm_connect(self.comm.world, self.comm.rank, handles)
      # (build() starts with a barrier: no rank scans before every rank has connected)

  def _check_zero_columns(self, total):
    if total > 0:                                     # giga.py:11-12, frankwolfe.py:11-12, orthopursuit.py:13-14
      raise ValueError(self.alg_name + '.__init__(): A must not have any 0 columns')

  # ---- state ---------------------------------------------------------------------------------
  def reset(self):
    self._native.reset()
    self.reached_numeric_limit = False

  def weights_sparse(self):
    """(global indices, weights) of the stored active rows with w > 0, ascending index order"""
    idx, w = self._native.active()
    keep = w > 0
    idx, w = idx[keep], w[keep]
    order = np.argsort(idx, kind='stable')
    return idx[order], w[order]

  def weights(self):
    w = np.zeros(self.n_global)
    idx, wa = self._native.active()
    w[idx] = wa
    return w

  @property
  def w(self):
    return self.weights()

  @w.setter
  def w(self, value):
    """assigning the dense weight vector (the reference's `self.w = ...`) becomes a sparse active-set write"""
    if self.comm.world > 1:
      raise NotImplementedError('assigning w is single-process only')
    value = np.asarray(value, dtype=np.float64)
    idx = np.flatnonzero(value != 0)
    self._native.set_active(idx, value[idx])

  def size(self):
    return int((self._native.active()[1] > 0).sum())

  def error(self):
    return self._native.error()

  # ---- greedy loop ---------------------------------------------------------------------------
  def _failure_message(self, e):
    if e.code == nat.IT_FAIL_CDIR:
      return 'cdirnrm < TOL: cdirnrm = ' + str(e.aux0)
    if e.code == nat.IT_FAIL_GAMMA:
      return 'precision loss in gammanum/gammadenom: num = ' + str(e.aux0) + ' denom = ' + str(e.aux1)
    if e.code == nat.IT_FAIL_MONOTONE:
      return 'Error not monotone: curr error = ' + str(e.aux0) + ' prev error = ' + str(e.aux1)
    return ''

  def _log_events(self, events):
    retried = False
    for e in events:
      if e.code == nat.IT_OK:
        retried = False
        continue
      self.log.warning('numerical precision error: ' + self._failure_message(e))
      if retried:
        self.log.warning('iterative step failed a second time. Assuming numeric limit reached.')
      else:
        self.log.warning('iterative step failed. Stabilizing and retrying...')
        retried = True

  def build(self, itrs):
    if self.reached_numeric_limit:
      self.log.warning('the numeric limit was already reached; returning. size = ' + str(self.size()) +
                       ', error = ' + str(self.error()))
      return
    if self.n_global*self._vecs.shape[1] == 0:
      self.log.warning('there are no data, returning.')
      return
    if self.comm.world > 1:
      self.comm.barrier()        # the ranks' solvers share one mailbox per context: no rank may still be in another solver's build
    self.last_events = self._run(int(itrs))
    self._log_events(self.last_events)
    if self.reached_numeric_limit:
      self.log.warning('the numeric limit has been reached. No more points will be added. size = ' +
                       str(self.size()) + ', error = ' + str(self.error()))

  def _run(self, itrs):
    events = self._native.build(itrs, util.TOL)
    self.reached_numeric_limit = self._native.halted()
    return events

  # ---- NNLS re-solve on the active set (snnls.py:82-97) --------------------------------------
  def optimize(self):
    prev_cost = self.error()
    idx, w = self._native.active()
    if not (w > 0).any():
      return
    self._native.nnls(from_scratch=True)            # float64 Lawson-Hanson on the device (csrc/nnls_logic.h)
    new_cost = self.error()
    if new_cost > prev_cost*(1. + util.TOL):
      self.log.warning('self.optimize() returned a solution with increasing error. Numeric limit possibly '
                       'reached: preverr = ' + str(prev_cost) + ' err = ' + str(new_cost) + '.')
      self._native.set_weights(w)
      self.reached_numeric_limit = True

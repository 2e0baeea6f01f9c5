"""Coreset API shell, kept verbatim in behaviour (reference: coreset/coreset.py:8-64):
incremental build(itrs), get() -> (wts, pts, idcs) with wts > 0, size(), optimize() with
revert-on-worse, reset()."""
import logging
import secrets
import numpy as np
from .. import util
from ..util import NumericalPrecisionError


class Coreset(object):
  def __init__(self):
    self.alg_name = self.__class__.__name__ + '-' + secrets.token_hex(3)
    self.log = logging.LoggerAdapter(logging.getLogger(), {'id': self.alg_name})
    self.reached_numeric_limit = False
    self._clear()

  def _clear(self):
    self.wts = np.array([])
    self.idcs = np.array([], dtype=np.int64)
    self.pts = np.array([])

  def reset(self):
    self._clear()
    self.reached_numeric_limit = False

  def size(self):
    return (self.wts > 0).sum()

  def get(self):
    if self.wts.shape[0] == 0:
      return np.array([]), np.array([]), np.array([])
    keep = self.wts > 0
    return self.wts[keep], self.pts[keep, :], self.idcs[keep]

  def error(self):
    raise NotImplementedError()

  def build(self, itrs):
    if self.reached_numeric_limit or itrs <= 0:
      return
    self._build(itrs)
    if self.reached_numeric_limit:
      self.log.warning('the numeric limit has been reached. No more points will be added. size = ' +
                       str(self.size()) + ', error = ' + str(self.error()))

  def optimize(self):
    saved = (self.wts.copy(), self.idcs.copy(), self.pts.copy())
    prev_cost = self.error()
    try:
      self._optimize()
      new_cost = self.error()
      if new_cost > prev_cost*(1. + util.TOL):
        raise NumericalPrecisionError('self.optimize() returned a solution with increasing error. Numeric limit '
                                      'possibly reached: preverr = ' + str(prev_cost) + ' err = ' + str(new_cost) + '.')
    except NumericalPrecisionError as e:
      self.log.warning(e)
      self.wts, self.idcs, self.pts = saved
      self.reached_numeric_limit = True

  def _optimize(self):
    raise NotImplementedError

  def _build(self, itrs):
    raise NotImplementedError

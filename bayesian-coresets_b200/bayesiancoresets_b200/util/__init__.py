"""Conventions shared with the reference's bayesiancoresets/util (util/__init__.py:1-7,
util/errors.py:1, util/log.py:5-7): a global tolerance, the NumericalPrecisionError type and a
verbosity switch on the root logger.  `nn_opt` is the host-side projected optimiser of
util/opt.py:4-28 (K-sized vectors; stays in Python)."""
import logging
import sys
import numpy as np

TOL = 1e-12


def set_tolerance(tol):
  global TOL
  TOL = tol


class NumericalPrecisionError(Exception):
  pass


_LEVELS = {'error': logging.ERROR, 'warning': logging.WARNING, 'critical': logging.CRITICAL, 'info': logging.INFO,
           'debug': logging.DEBUG, 'notset': logging.NOTSET}


def set_verbosity(verb):
  logging.getLogger().setLevel(_LEVELS[verb])


def _install_handler():
  root = logging.getLogger()
  if any(getattr(h, '_bcb200', False) for h in root.handlers):
    return
  h = logging.StreamHandler(sys.stderr)
  h._bcb200 = True

  class _Fmt(logging.Formatter):
    def format(self, record):
      if not hasattr(record, 'id'):
        record.id = record.name
      return super().format(record)
  h.setFormatter(_Fmt('%(levelname)s - %(id)s.%(funcName)s(): %(message)s'))
  root.addHandler(h)
  root.setLevel(logging.ERROR)


_install_handler()


def nn_opt(x0, grd, nn_idcs=None, opt_itrs=1000, step_sched=lambda i: 1./(i+1), b1=0.9, b2=0.999, eps=1e-8,
           verbose=False):
  """Bias-corrected first/second-moment steps with projection onto x >= 0 (util/opt.py:4-28)."""
  x = x0.copy()
  mom1 = np.zeros(x.shape[0])
  mom2 = np.zeros(x.shape[0])
  for i in range(opt_itrs):
    g = grd(x)
    mom1 = b1*mom1 + (1.-b1)*g
    mom2 = b2*mom2 + (1.-b2)*g**2
    x -= step_sched(i)*mom1/(1.-b1**(i+1))/(eps + np.sqrt(mom2/(1.-b2**(i+1))))
    if nn_idcs is None:
      x = np.maximum(x, 0.)
    else:
      x[nn_idcs] = np.maximum(x[nn_idcs], 0.)
    if verbose:
      sys.stdout.write('itr %d/%d\r' % (i+1, opt_itrs))
  return x

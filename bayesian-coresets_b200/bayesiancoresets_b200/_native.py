"""ctypes binding of the C-ABI shared library (include/bcg.h -> lib/libbcg_b200.so).

There is deliberately NO fallback: if the library is missing, or no CUDA device is visible,
every constructor raises.  The CPU oracle under /oracle is test infrastructure and is never
imported from here.
"""
import ctypes
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libbcg_b200.so')

BCG_OK = 0
ERR_NAMES = {1: 'BCG_ERR_CUDA', 2: 'BCG_ERR_ARG', 3: 'BCG_ERR_NO_DEVICE', 4: 'BCG_ERR_ZERO_B', 5: 'BCG_ERR_STATE',
             6: 'BCG_ERR_COMM', 7: 'BCG_ERR_UNSUPPORTED'}
ERR_ZERO_B = 4
ALG_GIGA, ALG_FW, ALG_OMP = 0, 1, 2
MODEL_LR, MODEL_GAUSSIAN, MODEL_POISSON = 0, 1, 2
IT_OK, IT_FAIL_CDIR, IT_FAIL_GEODESIC, IT_FAIL_GAMMA, IT_FAIL_MONOTONE = 0, 1, 2, 3, 4


class BcgError(RuntimeError):
  def __init__(self, code, msg):
    super().__init__('%s: %s' % (ERR_NAMES.get(code, 'BCG_ERR_%d' % code), msg))
    self.code = code


class IterEvent(ctypes.Structure):
  """struct bcg_iter_event"""
  _fields_ = [('code', ctypes.c_int32), ('nact', ctypes.c_int32), ('f', ctypes.c_int64), ('error', ctypes.c_double),
              ('aux0', ctypes.c_double), ('aux1', ctypes.c_double)]


_c = ctypes
_P = ctypes.c_void_p
_PP = ctypes.POINTER(ctypes.c_void_p)
_PROTOTYPES = {
  'bcg_abi_version': (_c.c_int, []),
  'bcg_last_error': (_c.c_char_p, []),
  'bcg_device_count': (_c.c_int, [_c.POINTER(_c.c_int)]),
  'bcg_ctx_create': (_c.c_int, [_c.c_int, _PP]),
  'bcg_ctx_destroy': (_c.c_int, [_P]),
  'bcg_ctx_info': (_c.c_int, [_P, _c.c_char_p, _c.c_int, _c.POINTER(_c.c_int), _c.POINTER(_c.c_int),
                              _c.POINTER(_c.c_int), _c.POINTER(_c.c_int64)]),
  'bcg_ctx_synchronize': (_c.c_int, [_P]),
  'bcg_ctx_trim': (_c.c_int, [_P]),
  'bcg_ctx_mem_info': (_c.c_int, [_P, _c.POINTER(_c.c_int64), _c.POINTER(_c.c_int64)]),
  'bcg_ctx_flush_l2': (_c.c_int, [_P, _c.c_int64]),
  'bcg_host_alloc': (_c.c_int, [_c.c_int64, _PP]),
  'bcg_host_free': (_c.c_int, [_P]),
  'bcg_vecs_from_host_f64': (_c.c_int, [_P, _P, _c.c_int64, _c.c_int32, _c.c_int64, _PP]),
  'bcg_vecs_project_lr': (_c.c_int, [_P, _P, _c.c_int64, _c.c_int32, _P, _c.c_int32, _PP]),
  'bcg_vecs_project_gaussian': (_c.c_int, [_P, _P, _c.c_int64, _c.c_int32, _P, _c.c_int32, _P, _PP]),
  'bcg_vecs_project_poisson': (_c.c_int, [_P, _P, _c.c_int64, _c.c_int32, _P, _c.c_int32, _PP]),
  'bcg_vecs_project_linear': (_c.c_int, [_P, _P, _c.c_int64, _c.c_int32, _P, _P, _c.c_int32, _PP]),
  'bcg_dataset_create': (_c.c_int, [_P, _P, _c.c_int64, _c.c_int32, _PP]),
  'bcg_dataset_destroy': (_c.c_int, [_P]),
  'bcg_dataset_project': (_c.c_int, [_P, _P, _c.c_int64, _c.c_int32, _c.c_int32, _P, _c.c_int32, _P, _PP, _P, _P]),
  'bcg_dataset_project_linear': (_c.c_int, [_P, _P, _c.c_int64, _c.c_int32, _P, _P, _c.c_int32, _PP, _P, _P]),
  'bcg_dataset_project_lazy': (_c.c_int, [_P, _c.c_int32, _c.c_int32, _P, _c.c_int32, _P, _PP]),
  'bcg_sampler_gaussian_post': (_c.c_int, [_P, _c.c_int32, _P, _P, _P, _P, _P, _c.c_int64, _P, _c.c_int32, _P, _P, _P]),
  'bcg_glm_joint': (_c.c_int, [_P, _c.c_int32, _P, _P, _c.c_int64, _c.c_int32, _c.c_int32, _P, _P, _P, _P]),
  'bcg_pseudo_grad': (_c.c_int, [_P, _c.c_int32, _P, _c.c_int64, _c.c_int32, _c.c_int32, _P, _c.c_int32, _P, _P, _P, _P, _P]),
  'bcg_dataset_audit': (_c.c_int, [_P, _c.c_int32, _c.c_int32, _P, _c.c_int32, _P, _c.c_int32, _P, _P, _P, _P]),
  'bcg_vecs_shape': (_c.c_int, [_P, _c.POINTER(_c.c_int64), _c.POINTER(_c.c_int32), _c.POINTER(_c.c_int32)]),
  'bcg_vecs_colsum': (_c.c_int, [_P, _P]),
  'bcg_vecs_norm_sum': (_c.c_int, [_P, _c.POINTER(_c.c_double)]),
  'bcg_vecs_zero_rows': (_c.c_int, [_P, _c.POINTER(_c.c_int64)]),
  'bcg_vecs_norms': (_c.c_int, [_P, _c.c_int64, _c.c_int64, _P]),
  'bcg_vecs_rows_f64': (_c.c_int, [_P, _c.c_int64, _c.c_int64, _P]),
  'bcg_vecs_destroy': (_c.c_int, [_P]),
  'bcg_solver_create': (_c.c_int, [_P, _P, _c.c_int32, _P, _c.c_double, _c.c_int64, _c.c_int64, _PP]),
  'bcg_solver_destroy': (_c.c_int, [_P]),
  'bcg_ctx_comm_handle': (_c.c_int, [_P, _P]),
  'bcg_solver_comm_handle': (_c.c_int, [_P, _P]),
  'bcg_solver_comm_connect': (_c.c_int, [_P, _c.c_int32, _c.c_int32, _P]),
  'bcg_solver_build': (_c.c_int, [_P, _c.c_int32, _c.c_double, _c.POINTER(IterEvent), _c.POINTER(_c.c_int32)]),
  'bcg_solver_omp_select': (_c.c_int, [_P, _c.POINTER(_c.c_int64)]),
  'bcg_solver_probe_argmax': (_c.c_int, [_P, _P, _c.POINTER(_c.c_int64), _c.POINTER(_c.c_double)]),
  'bcg_solver_nnls': (_c.c_int, [_P, _c.c_int32]),
  'bcg_solver_error': (_c.c_int, [_P, _c.POINTER(_c.c_double)]),
  'bcg_solver_size': (_c.c_int, [_P, _c.POINTER(_c.c_int64), _c.POINTER(_c.c_int64)]),
  'bcg_solver_halted': (_c.c_int, [_P, _c.POINTER(_c.c_int32)]),
  'bcg_solver_active': (_c.c_int, [_P, _c.c_int64, _P, _P, _c.POINTER(_c.c_int64)]),
  'bcg_solver_active_rows': (_c.c_int, [_P, _c.c_int64, _c.c_int64, _P]),
  'bcg_solver_set_weights': (_c.c_int, [_P, _P, _c.c_int64]),
  'bcg_solver_set_active': (_c.c_int, [_P, _P, _P, _c.c_int64]),
  'bcg_solver_reset': (_c.c_int, [_P]),
  'bcg_solver_timing': (_c.c_int, [_P, _c.POINTER(_c.c_float), _c.POINTER(_c.c_float), _c.POINTER(_c.c_int32),
                                   _c.POINTER(_c.c_int32)]),
  'bcg_solver_set_check_monotone': (_c.c_int, [_P, _c.c_int32]),
  'bcg_solver_set_force_exact': (_c.c_int, [_P, _c.c_int32]),
  'bcg_solver_exact_count': (_c.c_int, [_P, _c.POINTER(_c.c_int64)]),
  'bcg_solver_set_filter16': (_c.c_int, [_P, _c.c_int32]),
  'bcg_solver_filter16_stats': (_c.c_int, [_P, _c.POINTER(_c.c_int32), _c.POINTER(_c.c_int64)]),
  'bcg_solver_set_profiling': (_c.c_int, [_P, _c.c_int32]),
  'bcg_solver_set_trace': (_c.c_int, [_P, _c.c_int32]),
  'bcg_solver_get_trace': (_c.c_int, [_P, _c.c_int32, _P, _c.POINTER(_c.c_int32)]),
  'bcg_comm_create': (_c.c_int, [_c.c_char_p, _c.c_int32, _c.c_int32, _c.c_int32, _c.c_int32, _PP]),
  'bcg_comm_destroy': (_c.c_int, [_P]),
  'bcg_comm_rank': (_c.c_int, [_P, _c.POINTER(_c.c_int32), _c.POINTER(_c.c_int32)]),
  'bcg_comm_allgather': (_c.c_int, [_P, _P, _c.c_int64, _P]),
  'bcg_comm_allreduce_f64': (_c.c_int, [_P, _P, _c.c_int64, _c.c_int32]),
  'bcg_comm_barrier': (_c.c_int, [_P]),
}
EXPORTED_SYMBOLS = sorted(_PROTOTYPES)

_lib = None


def lib():
  """Load the shared library (once).  Raises if it has not been built."""
  global _lib
  if _lib is None:
    if not os.path.exists(LIB_PATH):
      raise ImportError('%s is missing: build it with `make -C bayesian-coresets_b200` (nvcc, sm_100a). '
                        'There is no CPU fallback.' % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOTYPES.items():
      fn = getattr(L, name)
      fn.restype = res
      fn.argtypes = args
    if L.bcg_abi_version() != 1:
      raise ImportError('libbcg_b200.so ABI version mismatch')
    _lib = L
  return _lib


def check(rc):
  if rc != BCG_OK:
    raise BcgError(rc, lib().bcg_last_error().decode('utf-8', 'replace'))


def _f64(a):
  return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
  return ctypes.c_void_p(a.ctypes.data)


class _PinnedBlock(object):
  """owner of one bcg_host_alloc block; freed when the last ndarray view of it is collected"""
  def __init__(self, nbytes):
    self.ptr = ctypes.c_void_p()
    check(lib().bcg_host_alloc(int(nbytes), ctypes.byref(self.ptr)))
    self.nbytes = int(nbytes)

  def __del__(self):
    try:
      if self.ptr:
        lib().bcg_host_free(self.ptr)
        self.ptr = None
    except Exception:
      pass


def pinned_empty(shape, dtype=np.float64):
  """ndarray in page-locked host memory (bcg_host_alloc): uploads from it are DMA copies straight from the array,
  without the staging memcpy a pageable ndarray needs.  Use it for the data handed to the coreset constructors."""
  dtype = np.dtype(dtype)
  shape = (shape,) if np.isscalar(shape) else tuple(shape)
  nbytes = int(np.prod(shape, dtype=np.int64))*dtype.itemsize
  blk = _PinnedBlock(max(nbytes, 1))
  buf = (ctypes.c_char*max(nbytes, 1)).from_address(blk.ptr.value)
  buf._bcg_owner = blk                                   # the ctypes buffer (base of the ndarray) keeps the block alive
  return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape, dtype=np.int64))).reshape(shape)


def pinned_copy(a):
  out = pinned_empty(np.shape(a), np.asarray(a).dtype)
  out[...] = a
  return out


class Context(object):
  """bcg_ctx: one CUDA device + private stream.  Fails loudly when no B200 is visible."""
  _default = {}

  def __init__(self, device=0):
    self.handle = ctypes.c_void_p()
    check(lib().bcg_ctx_create(int(device), ctypes.byref(self.handle)))
    self.device = int(device)

  @classmethod
  def default(cls, device=None):
    if device is None:
      device = int(os.environ.get('LOCAL_RANK', '0'))
    if device not in cls._default:
      cls._default[device] = cls(device)
    return cls._default[device]

  def info(self):
    name = ctypes.create_string_buffer(256)
    sm, maj, mnr, mem = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int64()
    check(lib().bcg_ctx_info(self.handle, name, 256, ctypes.byref(sm), ctypes.byref(maj), ctypes.byref(mnr),
                             ctypes.byref(mem)))
    return {'name': name.value.decode(), 'sm_count': sm.value, 'cc': (maj.value, mnr.value), 'total_mem': mem.value}

  def synchronize(self):
    check(lib().bcg_ctx_synchronize(self.handle))

  def trim(self):
    """release the cached matrix buffers of destroyed DeviceVecs"""
    check(lib().bcg_ctx_trim(self.handle))

  def mem_info(self):
    f, t = ctypes.c_int64(), ctypes.c_int64()
    check(lib().bcg_ctx_mem_info(self.handle, ctypes.byref(f), ctypes.byref(t)))
    return f.value, t.value

  def comm_handle(self):
    """64-byte IPC handle of this context's N-sharding mailbox"""
    buf = ctypes.create_string_buffer(64)
    check(lib().bcg_ctx_comm_handle(self.handle, buf))
    return buf.raw

  def flush_l2(self, nbytes=256 << 20):
    check(lib().bcg_ctx_flush_l2(self.handle, int(nbytes)))


def gaussian_post_sample(th0, Sig0inv, Siginv, pts, wts, E, want_post=False, ctx=None):
  """device evaluation of theta = mup + E U^T, (mup, U) = weighted_post(...) (bcg_sampler_gaussian_post)"""
  ctx = ctx or Context.default()
  th0, S0, Si, E = _f64(th0), _f64(Sig0inv), _f64(Siginv), _f64(np.atleast_2d(E))
  d = th0.shape[0]
  pts = _f64(np.reshape(pts, (-1, d))) if pts is not None and np.size(pts) else np.zeros((0, d))
  wts = _f64(np.reshape(wts, (-1,))) if wts is not None and np.size(wts) else np.zeros(0)
  theta = np.empty((E.shape[0], d))
  mup = np.empty(d) if want_post else None
  U = np.empty((d, d)) if want_post else None
  p = lambda x: None if x is None or x.size == 0 else _ptr(x)
  check(lib().bcg_sampler_gaussian_post(ctx.handle, d, _ptr(th0), _ptr(S0), _ptr(Si), p(pts), p(wts), pts.shape[0], p(E), E.shape[0],
                                        p(theta), p(mup), p(U)))
  return (theta, mup, U) if want_post else theta


def glm_joint(model, Z, w, theta, hess=False, ctx=None):
  """device evaluation of (log_joint, grad_th_log_joint[, hess_th_log_joint]) over the rows of Z (bcg_glm_joint)"""
  ctx = ctx or Context.default()
  theta = _f64(np.reshape(theta, (-1,)))
  d = theta.shape[0]
  Z = _f64(np.atleast_2d(Z)) if np.size(Z) else np.zeros((0, d + (1 if model == MODEL_POISSON else 0)))
  w = _f64(np.reshape(w, (-1,)))
  v = ctypes.c_double()
  g = np.empty(d)
  H = np.empty((d, d)) if hess else None
  check(lib().bcg_glm_joint(ctx.handle, model, _ptr(Z) if Z.size else None, _ptr(w) if w.size else None, Z.shape[0], Z.shape[1], d,
                            _ptr(theta), ctypes.byref(v), _ptr(g), None if H is None else _ptr(H)))
  return (v.value, g, H) if hess else (v.value, g)


def pseudo_grad(model, pts, theta, Siginv=None, w=None, resid=None, full=False, ctx=None):
  """device evaluation of the pseudo-point gradients (bcg_pseudo_grad): returns (glls (K, S, dz) | None, ugrad (K, dz) | None);
  ugrad is computed when w and resid are given"""
  ctx = ctx or Context.default()
  pts, theta = _f64(np.atleast_2d(pts)), _f64(np.atleast_2d(theta))
  S, d = theta.shape
  K, zld = pts.shape
  dz = d + 1 if model == MODEL_POISSON else d
  si = None if Siginv is None else _f64(Siginv)
  g = np.empty((K, S, dz)) if full else None
  u = np.empty((K, dz)) if w is not None else None
  ww = None if w is None else _f64(w)
  rr = None if resid is None else _f64(resid)
  p = lambda x: None if x is None else _ptr(x)
  check(lib().bcg_pseudo_grad(ctx.handle, model, _ptr(pts), K, zld, d, _ptr(theta), S, p(si), p(ww), p(rr), p(g), p(u)))
  return g, u


class Dataset(object):
  """bcg_dataset: the (n, d) data uploaded once and projected many times."""
  def __init__(self, Z, ctx=None):
    self.ctx = ctx or Context.default()
    Z = _f64(np.atleast_2d(Z))
    self.shape = Z.shape
    self.handle = ctypes.c_void_p()
    check(lib().bcg_dataset_create(self.ctx.handle, _ptr(Z), Z.shape[0], Z.shape[1], ctypes.byref(self.handle)))

  def project(self, model, theta, Siginv=None, vecs=False, rows=False, colsum=False, sub=None):
    """evaluate + row-centre on the device; returns (DeviceVecs | None, ndarray rows | None, ndarray colsum | None).
    sub: optional int64 row indices -- only those rows are projected (gathered on the device)."""
    theta = _f64(np.atleast_2d(theta))
    S, d = theta.shape
    si = None if Siginv is None else _f64(Siginv)
    hv = ctypes.c_void_p()
    idx = None if sub is None else np.ascontiguousarray(sub, dtype=np.int64)
    nsel = 0 if idx is None else idx.shape[0]
    sel = (None if idx is None else _ptr(idx), nsel)
    out_rows = np.empty((self.shape[0] if idx is None else nsel, S)) if rows else None
    out_cs = np.empty(S) if colsum else None
    outs = (ctypes.byref(hv) if vecs else None, None if out_rows is None else _ptr(out_rows),
            None if out_cs is None else _ptr(out_cs))
    if model == MODEL_GAUSSIAN:
      # after row-centring only x.(Siginv theta_s) - 0.5 theta_s.Siginv.theta_s survives (model_gaussian.py:4-10)
      A = _f64(theta.dot(si))
      coff = _f64(-0.5*(A*theta).sum(axis=1))
      check(lib().bcg_dataset_project_linear(self.handle, sel[0], sel[1], d, _ptr(A), _ptr(coff), S, *outs))
    else:
      check(lib().bcg_dataset_project(self.handle, sel[0], sel[1], model, d, _ptr(theta), S, None, *outs))
    return (DeviceVecs(self.ctx, hv) if vecs else None), out_rows, out_cs

  def project_lazy(self, model, theta, Siginv=None):
    """the never-materialising projection: a DeviceVecs without a matrix (norms / column sums only); solvers built over it
    re-evaluate the rows from this dataset at every selection pass"""
    theta = _f64(np.atleast_2d(theta))
    S, d = theta.shape
    si = None if Siginv is None else _f64(Siginv)
    hv = ctypes.c_void_p()
    check(lib().bcg_dataset_project_lazy(self.handle, model, d, _ptr(theta), S, None if si is None else _ptr(si), ctypes.byref(hv)))
    v = DeviceVecs(self.ctx, hv)
    v._source = self                                   # the dataset must outlive the lazy projection
    return v

  def audit(self, model, theta, Siginv=None, kind=ALG_FW, dirs=None, scores=True, norms=False, colsum=False):
    """independent float64 re-evaluation of one selection pass from the raw data (bcg_dataset_audit):
    returns (scores | None, norms | None, colsum | None) as host ndarrays"""
    theta = _f64(np.atleast_2d(theta))
    S, d = theta.shape
    si = None if Siginv is None else _f64(Siginv)
    dd = None if dirs is None else _f64(np.atleast_2d(dirs))
    n = self.shape[0]
    o_sc = np.empty(n) if (scores and dd is not None) else None
    o_nr = np.empty(n) if norms else None
    o_cs = np.empty(S) if colsum else None
    p = lambda x: None if x is None else _ptr(x)
    check(lib().bcg_dataset_audit(self.handle, model, d, _ptr(theta), S, p(si), kind, p(dd), p(o_sc), p(o_nr), p(o_cs)))
    return o_sc, o_nr, o_cs

  def __del__(self):
    try:
      if self.handle:
        lib().bcg_dataset_destroy(self.handle)
        self.handle = None
    except Exception:
      pass


class DeviceVecs(object):
  """bcg_vecs: device-resident (n, S) projection -- unit float32 rows + float64 norms.

  Stands in for the ndarray returned by Projector.project in the reference: supports `.shape`,
  `.T`, `.sum(axis=0)` (the three things coreset/hilbert.py:24 uses) and `.to_numpy()`."""

  def __init__(self, ctx, handle):
    self.ctx = ctx
    self.handle = handle
    n, S, ld = ctypes.c_int64(), ctypes.c_int32(), ctypes.c_int32()
    check(lib().bcg_vecs_shape(handle, ctypes.byref(n), ctypes.byref(S), ctypes.byref(ld)))
    self.shape = (n.value, S.value)
    self.ld = ld.value
    self.ndim = 2
    self.size = n.value*S.value

  # ---- constructors -----------------------------------------------------------------------
  @classmethod
  def from_host(cls, rows, ctx=None):
    ctx = ctx or Context.default()
    rows = np.asarray(rows)
    if rows.ndim != 2:
      raise ValueError('expected an (n, S) array')
    if rows.dtype != np.float64 or rows.strides[1] != 8 or rows.strides[0] % 8 != 0 or rows.strides[0] < 8*rows.shape[1]:
      rows = _f64(rows)
    h = ctypes.c_void_p()
    ld_host = rows.strides[0]//8 if rows.shape[0] > 1 else rows.shape[1]   # strides of <= 1 row are arbitrary
    check(lib().bcg_vecs_from_host_f64(ctx.handle, _ptr(rows), rows.shape[0], rows.shape[1], ld_host,
                                       ctypes.byref(h)))
    return cls(ctx, h)

  @classmethod
  def project_host(cls, model, Z, theta, Siginv=None, ctx=None):
    """project a HOST array once (upload pipelined with the kernels; the data are not kept on the device)"""
    ctx = ctx or Context.default()
    Z, theta = _f64(np.atleast_2d(Z)), _f64(np.atleast_2d(theta))
    S, d = theta.shape
    h = ctypes.c_void_p()
    if model == MODEL_GAUSSIAN:
      A = _f64(theta.dot(_f64(Siginv)))
      coff = _f64(-0.5*(A*theta).sum(axis=1))
      if Z.shape[1] != d:
        raise ValueError('x and theta disagree on the feature dimension')
      check(lib().bcg_vecs_project_linear(ctx.handle, _ptr(Z), Z.shape[0], d, _ptr(A), _ptr(coff), S, ctypes.byref(h)))
    elif model == MODEL_LR:
      if Z.shape[1] != d:
        raise ValueError('Z and theta disagree on the feature dimension')
      check(lib().bcg_vecs_project_lr(ctx.handle, _ptr(Z), Z.shape[0], d, _ptr(theta), S, ctypes.byref(h)))
    else:
      if Z.shape[1] != d + 1:
        raise ValueError('Z must be [x, y] with one more column than theta')
      check(lib().bcg_vecs_project_poisson(ctx.handle, _ptr(Z), Z.shape[0], d, _ptr(theta), S, ctypes.byref(h)))
    return cls(ctx, h)

  @classmethod
  def project_lr(cls, Z, theta, ctx=None):
    return cls.project_host(MODEL_LR, Z, theta, ctx=ctx)

  @classmethod
  def project_gaussian(cls, x, theta, Siginv, ctx=None):
    return cls.project_host(MODEL_GAUSSIAN, x, theta, Siginv, ctx=ctx)

  @classmethod
  def project_poisson(cls, Z, theta, ctx=None):
    return cls.project_host(MODEL_POISSON, Z, theta, ctx=ctx)

  # ---- ndarray-like surface -----------------------------------------------------------------
  @property
  def T(self):
    return DeviceVecsT(self)

  def sum(self, axis=None):
    if axis != 0:
      raise NotImplementedError('DeviceVecs.sum supports axis=0 only (hilbert.py:24)')
    out = np.empty(self.shape[1])
    check(lib().bcg_vecs_colsum(self.handle, _ptr(out)))
    return out

  def norm_sum(self):
    v = ctypes.c_double()
    check(lib().bcg_vecs_norm_sum(self.handle, ctypes.byref(v)))
    return v.value

  def zero_rows(self):
    v = ctypes.c_int64()
    check(lib().bcg_vecs_zero_rows(self.handle, ctypes.byref(v)))
    return v.value

  def norms(self, row0=0, nrows=None):
    nrows = self.shape[0] - row0 if nrows is None else nrows
    out = np.empty(nrows)
    check(lib().bcg_vecs_norms(self.handle, row0, nrows, _ptr(out)))
    return out

  def to_numpy(self, row0=0, nrows=None):
    """(nrows, S) float64 = norm * unit row; the host view of what the device holds."""
    nrows = self.shape[0] - row0 if nrows is None else nrows
    out = np.empty((nrows, self.shape[1]))
    check(lib().bcg_vecs_rows_f64(self.handle, row0, nrows, _ptr(out)))
    return out

  def argmax_dot(self, direction):
    """(row, value) of max_n <a_n/||a_n||, direction> over the rows (float64 re-scored, ties -> lowest row)"""
    if getattr(self, '_probe', None) is None:
      self._probe = NativeSolver(self, ALG_FW, np.ones(self.shape[1]), 0., hold=False)
    return self._probe.probe_argmax(direction)

  def __array__(self, dtype=None, copy=None):
    a = self.to_numpy()
    return a if dtype is None else a.astype(dtype)

  def __del__(self):
    try:
      self._probe = None
      if self.handle:
        lib().bcg_vecs_destroy(self.handle)
        self.handle = None
    except Exception:
      pass


class DeviceVecsT(object):
  """`vecs.T`: the (S, N) view the reference hands to the snnls constructors (hilbert.py:24)."""
  def __init__(self, vecs):
    self.vecs = vecs
    self.shape = (vecs.shape[1], vecs.shape[0])
    self.size = vecs.size

  @property
  def T(self):
    return self.vecs


class NativeSolver(object):
  """bcg_solver"""
  def __init__(self, vecs, alg, b, norm_sum, row_offset=0, n_global=None, hold=True):
    self.vecs = vecs if hold else None     # keep the matrix alive (not for a probe owned BY the matrix: no cycle)
    self._S = vecs.shape[1]
    b = _f64(b)
    if b.shape != (vecs.shape[1],):
      raise ValueError('b must have shape (S,)')
    self.handle = ctypes.c_void_p()
    n_global = vecs.shape[0] if n_global is None else n_global
    check(lib().bcg_solver_create(vecs.ctx.handle, vecs.handle, alg, _ptr(b), float(norm_sum), int(row_offset),
                                  int(n_global), ctypes.byref(self.handle)))

  def comm_handle(self):
    buf = ctypes.create_string_buffer(64)
    check(lib().bcg_solver_comm_handle(self.handle, buf))
    return buf.raw

  def comm_connect(self, world, rank, handles):
    blob = b''.join(handles)
    assert len(blob) == 64*world
    check(lib().bcg_solver_comm_connect(self.handle, world, rank, ctypes.c_char_p(blob)))

  def build(self, itrs, tol):
    ev = (IterEvent*max(int(itrs), 1))()
    n = ctypes.c_int32()
    check(lib().bcg_solver_build(self.handle, int(itrs), float(tol), ev, ctypes.byref(n)))
    return [ev[i] for i in range(n.value)]

  def omp_select(self):
    f = ctypes.c_int64()
    check(lib().bcg_solver_omp_select(self.handle, ctypes.byref(f)))
    return f.value

  def probe_argmax(self, direction):
    d = _f64(direction)
    f, sc = ctypes.c_int64(), ctypes.c_double()
    check(lib().bcg_solver_probe_argmax(self.handle, _ptr(d), ctypes.byref(f), ctypes.byref(sc)))
    return f.value, sc.value

  def nnls(self, from_scratch=False):
    check(lib().bcg_solver_nnls(self.handle, 1 if from_scratch else 0))

  def error(self):
    v = ctypes.c_double()
    check(lib().bcg_solver_error(self.handle, ctypes.byref(v)))
    return v.value

  def halted(self):
    v = ctypes.c_int32()
    check(lib().bcg_solver_halted(self.handle, ctypes.byref(v)))
    return bool(v.value)

  def active(self):
    k = ctypes.c_int64()
    check(lib().bcg_solver_active(self.handle, 0, None, None, ctypes.byref(k)))
    idx = np.empty(k.value, dtype=np.int64)
    w = np.empty(k.value)
    if k.value:
      check(lib().bcg_solver_active(self.handle, k.value, _ptr(idx), _ptr(w), ctypes.byref(k)))
    return idx, w

  def active_rows(self, first, count):
    out = np.empty((count, self._S))
    if count:
      check(lib().bcg_solver_active_rows(self.handle, first, count, _ptr(out)))
    return out

  def set_weights(self, w):
    w = _f64(w)
    check(lib().bcg_solver_set_weights(self.handle, _ptr(w), w.shape[0]))

  def set_active(self, idx, w):
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    w = _f64(w)
    check(lib().bcg_solver_set_active(self.handle, _ptr(idx), _ptr(w), idx.shape[0]))

  def reset(self):
    check(lib().bcg_solver_reset(self.handle))

  def set_check_monotone(self, on):
    check(lib().bcg_solver_set_check_monotone(self.handle, 1 if on else 0))

  def set_force_exact(self, on):
    check(lib().bcg_solver_set_force_exact(self.handle, 1 if on else 0))

  def exact_count(self):
    v = ctypes.c_int64()
    check(lib().bcg_solver_exact_count(self.handle, ctypes.byref(v)))
    return v.value

  def set_filter16(self, on):
    """float16 pre-filter of the persistent greedy kernels (bcg.h: bcg_solver_set_filter16); selections are unchanged"""
    check(lib().bcg_solver_set_filter16(self.handle, 1 if on else 0))

  def filter16_stats(self):
    """(enabled, rows re-scanned in float32 so far)"""
    e, r = ctypes.c_int32(), ctypes.c_int64()
    check(lib().bcg_solver_filter16_stats(self.handle, ctypes.byref(e), ctypes.byref(r)))
    return bool(e.value), int(r.value)

  def set_profiling(self, on):
    check(lib().bcg_solver_set_profiling(self.handle, 1 if on else 0))

  def set_trace(self, on):
    check(lib().bcg_solver_set_trace(self.handle, 1 if on else 0))

  def trace(self):
    """(iters, 8) uint64 device timestamps (globaltimer ns) of the last build: grid arrived, direction
    published, scan start, scan end (CTA 0 warp 0), then 4 control-warp milestones"""
    n = ctypes.c_int32()
    check(lib().bcg_solver_get_trace(self.handle, 0, None, ctypes.byref(n)))
    out = np.zeros((n.value, 8), dtype=np.uint64)
    if n.value:
      check(lib().bcg_solver_get_trace(self.handle, n.value, _ptr(out), ctypes.byref(n)))
    return out

  def timing(self):
    b, s = ctypes.c_float(), ctypes.c_float()
    ns, nt = ctypes.c_int32(), ctypes.c_int32()
    check(lib().bcg_solver_timing(self.handle, ctypes.byref(b), ctypes.byref(s), ctypes.byref(ns), ctypes.byref(nt)))
    return {'build_ms': b.value, 'scan_ms': s.value, 'scan_launches': ns.value, 'step_launches': nt.value}

  def __del__(self):
    try:
      if self.handle:
        lib().bcg_solver_destroy(self.handle)
        self.handle = None
    except Exception:
      pass

"""TEST DOUBLE for the native layer (used only by the gloo CPU tests of the host-side N-sharding
logic).  It mimics the device protocol in float64 NumPy: every rank scores its own shard, the
candidates (score, global index, row) are all-gathered -- the stand-in for the NVLink mailbox
exchange of csrc/step_kernels.cuh -- the winner is the maximum score with ties to the lowest
global index, and the GIGA reweight is replicated on every rank."""
import numpy as np


RANK = [0]      # set by the worker: the fake mailbox handle carries the rank


class FakeCtx(object):
  def comm_handle(self):
    return bytes([RANK[0]])*64


class FakeVecs(object):
  def __init__(self, rows):
    self.rows = np.array(rows, dtype=np.float64)
    self.shape = self.rows.shape
    self.size = self.rows.size
    self.ctx = FakeCtx()

  @classmethod
  def from_host(cls, rows, ctx=None):
    return cls(rows)

  @property
  def T(self):
    from bayesiancoresets_b200._native import DeviceVecsT
    return DeviceVecsT(self)

  def sum(self, axis=0):
    return self.rows.sum(axis=0)

  def norms(self):
    return np.sqrt((self.rows**2).sum(axis=1))

  def norm_sum(self):
    return float(self.norms().sum())

  def zero_rows(self):
    return int((self.norms() == 0).sum())


class _Ev(object):
  def __init__(self, code, f, error):
    self.code, self.f, self.error, self.aux0, self.aux1, self.nact = code, f, error, 0., 0., 0


def make_fake_solver(comm):
  class FakeSolver(object):
    def __init__(self, vecs, alg, b, norm_sum, row_offset=0, n_global=None):
      assert alg == 0, 'the test double implements GIGA only'
      self.A = vecs.rows
      self.norms = vecs.norms()
      self.An = self.A/self.norms[:, None] if self.A.shape[0] else self.A
      self.b = np.array(b)
      self.bnorm = np.sqrt((self.b**2).sum())
      self.bn = self.b/self.bnorm
      self.row_offset, self.n_global = row_offset, n_global
      self.idx, self.w, self.rows = [], [], []
      self.connected = comm.world == 1

    def comm_handle(self):
      return bytes([comm.rank])*64

    def comm_connect(self, world, rank, handles):
      assert world == comm.world and rank == comm.rank
      assert [h[0] for h in handles] == list(range(world))
      self.connected = True

    def _xw(self):
      xw = np.zeros(self.b.shape[0])
      for wk, r in zip(self.w, self.rows):
        xw += wk*r
      return xw

    def error(self):
      return float(np.sqrt(((self._xw() - self.b)**2).sum()))

    def halted(self):
      return False

    def active(self):
      return np.array(self.idx, dtype=np.int64), np.array(self.w, dtype=np.float64)

    def build(self, itrs, tol):
      assert self.connected
      events = []
      for _ in range(itrs):
        xw = self._xw()
        nw = np.sqrt((xw**2).sum())
        nw = 1. if nw == 0. else nw
        xwn = xw/nw
        cdir = self.bn - self.bn.dot(xwn)*xwn
        cdir /= np.sqrt((cdir**2).sum())
        if self.An.shape[0]:
          s0, s1 = self.An.dot(cdir), self.An.dot(xwn)
          ok = np.logical_and(s1 > -1. + 1e-14, 1. - s1**2 > 0.)
          den = np.where(ok, np.sqrt(np.where(ok, 1. - s1**2, 1.)), np.inf)
          sc = s0/den
          l = int(sc.argmax())
          cand = (float(sc[l]), self.row_offset + l, self.A[l].copy())
        else:
          cand = (-np.inf, -1, None)
        cands = [c for c in comm.allgather_object(cand) if c[1] >= 0]
        score, f, xf = max(cands, key=lambda c: (c[0], -c[1]))
        nf = np.sqrt((xf**2).sum())
        gA = self.bn.dot(xf/nf) - self.bn.dot(xwn)*xwn.dot(xf/nf)
        gB = self.bn.dot(xwn) - self.bn.dot(xf/nf)*xwn.dot(xf/nf)
        a, b = gB/(gA+gB)/nw, gA/(gA+gB)/nf
        x = a*xw + b*xf
        nx = np.sqrt((x**2).sum())
        scale = self.bnorm/nx*(x/nx).dot(self.bn)
        self.w = [a*scale*wk for wk in self.w]
        if f in self.idx:
          k = self.idx.index(f)
          self.w[k] = max(0., self.w[k] + b*scale)
        else:
          self.idx.append(f); self.w.append(max(0., b*scale)); self.rows.append(xf)
        events.append(_Ev(0, f, self.error()))
      return events
  return FakeSolver

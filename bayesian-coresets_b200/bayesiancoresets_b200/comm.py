"""Host-side plumbing for N-sharding over the GPUs of one node: one process per GPU, launched by
torchrun; torch.distributed (NCCL on the GPU box, gloo in CPU tests) is used ONLY for bootstrap
(exchanging 64-byte mailbox handles, row counts) and for the one-off reductions of b and
sum ||a_n|| after a projection.  The per-iteration exchange of the greedy loop never goes through
this module: it is fused into the step kernel over NVLink peer memory (csrc/step_kernels.cuh).
"""
import os
import numpy as np


class SerialComm(object):
  """world of one (the reference's situation)"""
  rank = 0
  world = 1

  def allreduce_sum(self, arr):
    return np.array(arr, dtype=np.float64, copy=True)

  def allgather_object(self, obj):
    return [obj]

  def barrier(self):
    pass


class TorchComm(object):
  """torch.distributed-backed communicator; the process group must already be initialised."""
  def __init__(self, group=None):
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
      raise RuntimeError('torch.distributed is not initialised')
    self._torch, self._dist, self.group = torch, dist, group
    self.rank = dist.get_rank(group)
    self.world = dist.get_world_size(group)
    self.backend = dist.get_backend(group)
    if self.backend == 'nccl':
      self.device = torch.device('cuda', int(os.environ.get('LOCAL_RANK', self.rank)))
      torch.cuda.set_device(self.device)
    else:
      self.device = torch.device('cpu')

  def allreduce_sum(self, arr):
    t = self._torch.from_numpy(np.array(arr, dtype=np.float64, copy=True)).to(self.device)
    self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM, group=self.group)
    return t.cpu().numpy()

  def allgather_object(self, obj):
    out = [None]*self.world
    self._dist.all_gather_object(out, obj, group=self.group)
    return out

  def barrier(self):
    self._dist.barrier(group=self.group)


def shard_layout(comm, n_local):
  """(row_offset of this rank, global row count, per-rank counts) for contiguous N-sharding."""
  counts = [int(c) for c in comm.allgather_object(int(n_local))]
  return sum(counts[:comm.rank]), sum(counts), counts


def even_shard(n_total, rank, world):
  """contiguous near-even split of range(n_total): returns (lo, hi) of `rank`"""
  base, rem = divmod(int(n_total), int(world))
  lo = rank*base + min(rank, rem)
  return lo, lo + base + (1 if rank < rem else 0)


def gather_rows(comm, data_local, row_offset, idcs_global):
  """rows `idcs_global` of the row-sharded dataset, assembled on every rank"""
  idcs_global = np.asarray(idcs_global, dtype=np.int64)
  mine = (idcs_global >= row_offset) & (idcs_global < row_offset + data_local.shape[0])
  piece = (np.flatnonzero(mine), data_local[idcs_global[mine] - row_offset])
  out = np.empty((idcs_global.shape[0],) + data_local.shape[1:], dtype=data_local.dtype)
  for pos, rows in comm.allgather_object(piece):
    out[pos] = rows
  return out

"""phase timing of the end-to-end job under N ranks: where do the occasional 100+ ms stalls of the first jobs come from?"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bayesian-coresets_b200')); sys.path.insert(0, ROOT)
import numpy as np
import bayesiancoresets_b200 as bc
from bayesiancoresets_b200 import _native as nat
from bench import lr_shard, lr_samples
N, d, S, steps = 10_000_000, 10, 512, 20
comm = bc.comm.default_comm()
rank, world = comm.rank, comm.world
ctx = bc.Context.default(int(os.environ.get('LOCAL_RANK', '0')))
lo, hi = bc.comm.even_shard(N, rank, world)
Z, th = lr_shard(0, lo, hi, d)
theta = lr_samples(0, th, S)
prj = bc.LogisticRegressionProjector(lambda n, w, p: theta, S, ctx=ctx)
kw = {'comm': comm} if world > 1 else {}
# the same prelude as bench.py: one resident solver first
from bench import ClockSampler
cs = bc.HilbertCoreset(Z, prj, **kw); cs.build(5)
smp = ClockSampler(int(os.environ.get('LOCAL_RANK', '0')))
if len(sys.argv) > 2 and sys.argv[2] == 'sampler' and rank == 0: smp.start()
cs.build(steps); ctx.synchronize(); comm.barrier()
if len(sys.argv) > 2 and sys.argv[2] == 'sampler' and rank == 0: print('clocks', smp.stop(), flush=True)
del cs
mode = sys.argv[1] if len(sys.argv) > 1 else 'pinned'
src = bc.pinned_copy(Z) if mode == 'pinned' else Z
for rep in range(6):
  comm.barrier()
  t = [time.perf_counter()]
  vecs = prj.project_device(src); ctx.synchronize(); t.append(time.perf_counter())
  b = vecs.sum(axis=0)
  if world > 1: b = comm.allreduce_sum(b)
  t.append(time.perf_counter())
  sol = bc.snnls.GIGA(vecs.T, b, **kw); t.append(time.perf_counter())
  sol.build(steps); ctx.synchronize(); t.append(time.perf_counter())
  idx, w = sol.weights_sparse(); err = sol.error(); t.append(time.perf_counter())
  comm.barrier(); t.append(time.perf_counter())
  ph = [round((t[i+1]-t[i])*1e3, 2) for i in range(len(t)-1)]
  for r in range(world):
    comm.barrier()
    if r == rank: print(json.dumps({'rep': rep, 'rank': rank, 'mode': mode, 'project': ph[0], 'allreduce_b': ph[1], 'solver_ctor': ph[2], 'build': ph[3], 'export': ph[4], 'barrier': ph[5], 'total': round((t[-1]-t[0])*1e3, 2), 'device_build_ms': round(sol._native.timing()['build_ms'], 2)}), flush=True)
  del sol, vecs
comm.barrier(); comm.close()

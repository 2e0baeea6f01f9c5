#!/usr/bin/env python
"""Secondary measurements for the BASELINE.json configs other than the headline:

  C2  LR HilbertCoreset GIGA / FW / OMP, N=1e6 S=256            (iters/s per solver)
  C3  Gaussian SparseVICoreset, N=1e6 d=200 S=512               (s per projector pass, s per build iteration)
  C4  LR HilbertCoreset OrthoPursuit, N=1e7 S=512 (N-sharded)   (iters/s, K = 200)
  C5  Poisson BatchPSVICoreset gradient, N=1e7 d=128 S=512      (s per gradient evaluation)

  python bench_configs.py --config c3 [--scale 0.1]      (torchrun for N-sharded runs)
Prints one JSON line per measurement.  --scale shrinks N (for quick functional runs).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'bayesian-coresets_b200'))
from bench import lr_shard, lr_samples  # noqa: E402


def emit(**kw):
  if int(os.environ.get('RANK', '0')) == 0:
    print(json.dumps(kw), flush=True)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--config', required=True, choices=['c2', 'c3', 'c4', 'c5'])
  ap.add_argument('--scale', type=float, default=1.0)
  ap.add_argument('--iters', type=int, default=0)
  args = ap.parse_args()
  import bayesiancoresets_b200 as bc
  rank = int(os.environ.get('RANK', '0'))
  world = int(os.environ.get('WORLD_SIZE', '1'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  comm = None
  if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    comm = bc.comm.TorchComm()
  ctx = bc.Context.default(local_rank)
  kw = {'comm': comm} if comm is not None else {}

  if args.config in ('c2', 'c4'):
    N, d, S = (1_000_000, 10, 256) if args.config == 'c2' else (10_000_000, 10, 512)
    N = int(N*args.scale)
    K = args.iters or 200
    lo, hi = bc.comm.even_shard(N, rank, world)
    Z, th_true = lr_shard(0, lo, hi, d)
    theta = lr_samples(0, th_true, S)
    prj = bc.LogisticRegressionProjector(lambda n, w, p: theta, S, ctx=ctx)
    algs = [('OrthoPursuit', bc.snnls.OrthoPursuit)] if args.config == 'c4' else \
           [('GIGA', bc.snnls.GIGA), ('FrankWolfe', bc.snnls.FrankWolfe), ('OrthoPursuit', bc.snnls.OrthoPursuit)]
    for name, cls in algs:
      t0 = time.perf_counter()
      cs = bc.HilbertCoreset(Z, prj, snnls=cls, **kw)
      t_setup = time.perf_counter() - t0
      cs.build(5)
      ctx.synchronize()
      t0 = time.perf_counter()
      cs.build(K)
      ctx.synchronize()
      dt = time.perf_counter() - t0
      emit(config=args.config, alg=name, N=N, S=S, world=world, iters=K, iters_per_s=K/dt, ms_per_iter=1e3*dt/K,
           setup_s=t_setup, size=int(cs.snnls.size()), error=cs.error(),
           roofline_frac=(4.*(hi - lo)*S*K/dt/1e9)/6542.4)

  if args.config == 'c3':
    N, d, S = int(1_000_000*args.scale), 200, 512
    rng = np.random.RandomState(0)
    x = rng.randn(N, d) + 1.                                   # examples/gaussian/main.py:72,82: N(1_d, I)
    th0, Sig0inv, Siginv = np.zeros(d), np.eye(d), np.eye(d)

    def sampler_w(n, wts, pts):                                # examples/gaussian/main.py:107-113
      if wts is None or pts is None or pts.shape[0] == 0:
        wts, pts = np.zeros(1), np.zeros((1, d))
      prec = Sig0inv + wts.sum()*Siginv
      cov = np.linalg.inv(prec)
      mu = cov.dot(Sig0inv.dot(th0) + Siginv.dot((wts[:, None]*pts).sum(axis=0)))
      return mu + np.random.randn(n, d).dot(np.linalg.cholesky(cov).T)
    prj = bc.GaussianProjector(sampler_w, S, Siginv, ctx=ctx)
    t0 = time.perf_counter()
    prj.project_sum(x)                                         # includes the one-off upload of x
    t_first = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(5):
      prj.project_sum(x)
    t_sum = (time.perf_counter() - t0)/5
    t0 = time.perf_counter()
    v = prj.project_device(x, cache=True)
    t_full = time.perf_counter() - t0
    emit(config='c3', what='projector pass', N=N, d=d, S=S, colsum_only_s=t_sum, materialising_s=t_full,
         first_pass_with_upload_s=t_first, gflops_f64=2.*N*d*S/t_sum/1e9)
    del v
    opt_itrs = 100
    svi = bc.SparseVICoreset(x, prj, opt_itrs=opt_itrs)
    svi.build(1)
    t0 = time.perf_counter()
    svi.build(args.iters or 3)
    dt = (time.perf_counter() - t0)/(args.iters or 3)
    emit(config='c3', what='SparseVI build iteration', N=N, d=d, S=S, opt_itrs=opt_itrs, s_per_build_iter=dt,
         size=int(svi.size()))
    # apples-to-apples FW iters/s on the same N x S
    cs = bc.HilbertCoreset(x, prj, snnls=bc.snnls.FrankWolfe)
    cs.build(5)
    t0 = time.perf_counter()
    cs.build(200)
    dt = time.perf_counter() - t0
    emit(config='c3', what='Hilbert FrankWolfe on the Gaussian projection', N=N, S=S, iters_per_s=200/dt)

  if args.config == 'c5':
    N, d, S, sz = int(10_000_000*args.scale), 128, 512, 100
    lo, hi = bc.comm.even_shard(N, rank, world)
    rng = np.random.RandomState(1 + rank)
    th_true = np.random.RandomState(0).randn(d)/np.sqrt(d)
    X = np.hstack((rng.randn(hi - lo, d - 1), np.ones((hi - lo, 1))))
    y = rng.poisson(np.log1p(np.exp(X.dot(th_true)))).astype(np.float64)
    Z = np.hstack((X, y[:, None]))
    theta = th_true + 0.05*np.random.RandomState(2).randn(S, d)
    prj = bc.PoissonProjector(lambda n, w, p: theta, S, ctx=ctx)
    bp = bc.BatchPSVICoreset(Z, prj, opt_itrs=1)
    x0 = np.hstack((np.full(sz, N/sz), Z[:sz].reshape(-1)))
    t0 = time.perf_counter()
    bp.gradient(x0.copy(), sz, d + 1)                          # includes the one-off upload of Z
    t_first = time.perf_counter() - t0
    reps = args.iters or 5
    t0 = time.perf_counter()
    for _ in range(reps):
      g = bp.gradient(x0.copy(), sz, d + 1)
    dt = (time.perf_counter() - t0)/reps
    emit(config='c5', what='BatchPSVI gradient evaluation (local shard; the S-vector all-reduce is not included)',
         N=N, rows_local=hi - lo, d=d, S=S, K=sz, world=world, s_per_grad=dt, first_with_upload_s=t_first,
         gflops_f64=2.*(hi - lo)*d*S/dt/1e9, grad_norm=float(np.linalg.norm(g)))

  if comm is not None:
    import torch.distributed as dist
    comm.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
  main()

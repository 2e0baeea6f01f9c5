#!/usr/bin/env python
"""Benchmark of the greedy sparse-NNLS hot path (BASELINE.json metric: greedy iters/sec on an
N x S projection).

  python bench.py --gpus 1 --steps K --warmup W            # this engine (CUDA, sm_100a)
  python bench.py --impl reference --steps K --warmup W    # CPU reference path (oracle port, NumPy/OpenBLAS)
  torchrun ... bench.py --gpus N ...                        # N-sharded over N GPUs of one node

Workload (config.workload): the north-star target -- synthetic logistic regression, N = 1e7
datapoints, d = 10, S = 512 posterior samples, GIGA.  A "step" is ONE greedy iteration (one full
pass of the scan kernel over the resident N x S matrix plus the reweight).  With --gpus N the
same N rows are sharded over the ranks (strong scaling): each rank scans N/ranks rows and the
candidates are exchanged over NVLink peer memory inside the step kernel.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'bayesian-coresets_b200'))

WORKLOADS = {
  # name: (N, d, S)
  'lr_giga_N1e7_S512': (10_000_000, 10, 512),     # north-star target / BASELINE configs[3] shape
  'lr_giga_N1e6_S256': (1_000_000, 10, 256),      # BASELINE configs[1]
  'lr_giga_N1e6_S512': (1_000_000, 10, 512),
  'lr_giga_N2e5_S256': (200_000, 10, 256),        # quick functional check
}
METRIC = 'greedy_iters_per_sec'
UNIT = 'iters/s'


def lr_shard(seed, lo, hi, d):
  """rows [lo, hi) of the synthetic LR dataset (recipe of examples/simple_lr/main.py:22-35): x ~ N(0, I),
  y = +-1 with P(y=1) = sigmoid(x.th*), z = y x.  Generated in blocks so any shard is reproducible."""
  rng = np.random.RandomState(seed)
  th_true = rng.randn(d)
  Z = np.empty((hi - lo, d))
  blk = 1_000_000
  for b0 in range((lo // blk) * blk, hi, blk):
    r = np.random.RandomState(seed * 7919 + 1 + b0 // blk)
    X = r.randn(blk, d)
    y = np.where(r.rand(blk) <= 1. / (1. + np.exp(-X.dot(th_true))), 1., -1.)
    a, b = max(lo, b0), min(hi, b0 + blk)
    Z[a - lo:b - lo] = (y[:, None] * X)[a - b0:b - b0]
  return Z, th_true


def lr_samples(seed, th_true, S):
  """stand-in for the (host-side, untimed) Laplace posterior sampler: th* + 0.1 N(0, I)"""
  return th_true + 0.1 * np.random.RandomState(seed + 12345).randn(S, th_true.shape[0])


class ClockSampler(object):
  """nvidia-smi clocks / throttle reasons sampled DURING the timed region"""
  Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
       'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
       'clocks_event_reasons.sw_power_cap')

  def __init__(self, device):
    self.rows, self.proc, self.device = [], None, device

  def start(self):
    try:
      self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.device), '--query-gpu=' + self.Q,
                                    '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE,
                                   stderr=subprocess.DEVNULL, text=True)
      self.t = threading.Thread(target=self._read, daemon=True)
      self.t.start()
    except Exception:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append([c.strip() for c in line.split(',')])

  def stop(self):
    if self.proc is None:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
    time.sleep(0.15)
    self.proc.terminate()
    self.t.join(timeout=2)
    sm, smax, reasons = [], [], set()
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    for r in self.rows:
      try:
        sm.append(float(r[1])); smax.append(float(r[2]))
        for nm, v in zip(names, r[4:8]):
          if v.lower().startswith('active'):
            reasons.add(nm)
      except Exception:
        pass
    return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(smax) if smax else None,
            'reasons': sorted(reasons), 'samples': len(sm)}


def measured_peak():
  try:
    with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
      return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
  except Exception:
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def measured_traffic_ratio(kernel, S):
  """DRAM bytes / algorithmic bytes of the dominant kernel from the committed `ncu --set full` capture
  (profiles/r01_loop_traffic.json, written by profiles/extract_traffic.py); None when there is no capture"""
  try:
    with open(os.path.join(ROOT, 'profiles', 'r01_loop_traffic.json')) as f:
      t = json.load(f)
    if t.get('kernel') == kernel and int(t.get('S', -1)) == int(S):
      return float(t['dram_bytes']) / float(t['algorithmic_bytes'])
  except Exception:
    pass
  return None


def cpu_reference_run(workload, steps, warmup, sample_rows=None):
  """The reference's CPU path for this workload: float64 NumPy/OpenBLAS port of
  HilbertCoreset(...GIGA).build (oracle/, bit-identical to the reference) on all host cores, on a
  bounded row sample; the loop is memory-bound and linear in N, so iters/s is scaled by sample/N."""
  from oracle import greedy, models
  N, d, S = WORKLOADS[workload]
  cores = os.cpu_count()
  try:
    # torchrun exports OMP_NUM_THREADS=1: give the CPU reference every host core it can use
    import threadpoolctl
    threadpoolctl.threadpool_limits(limits=cores)
    used = [p.get('num_threads') for p in threadpoolctl.threadpool_info() if p.get('user_api') == 'blas']
    cores = max(used) if used else cores
  except Exception:
    pass
  if sample_rows is None:
    sample_rows = min(N, max(20_000, int(1.28e8 / S)))      # ~1 GB float64 matrix (+1 GB copy)
  Z, th_true = lr_shard(0, 0, sample_rows, d)
  theta = lr_samples(0, th_true, S)
  vecs = models.project(models.lr_loglik, Z, theta)
  o = greedy.GigaOracle(vecs.T, vecs.sum(axis=0))
  o.build(max(warmup, 1))
  t0 = time.perf_counter()
  o.build(steps)
  dt = time.perf_counter() - t0
  its = steps / dt
  return {'value': its * sample_rows / N, 'unit': UNIT, 'cores': cores, 'kind': 'port',
          'sample': '%d of %d rows (S=%d, float64), %d timed GIGA iterations at %.3f s/iter on the sample; '
                    'iters/s scaled by sample/N' % (sample_rows, N, S, steps, dt / steps),
          'sample_iters_per_s': its}, dt


def run_reference(args):
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  steps = min(args.steps, 40)
  cb, dt = cpu_reference_run(args.workload, steps, min(args.warmup, 3))
  N, d, S = WORKLOADS[args.workload]
  line = {'impl': 'reference', 'metric': METRIC, 'value': cb['value'], 'unit': UNIT, 'n_gpus': args.gpus,
          'steps': steps, 'warmup': min(args.warmup, 3), 'ms_per_step': 1e3 / cb['value'],
          'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
          'config': {'workload': args.workload, 'N': N, 'd': d, 'S': S, 'alg': 'GIGA'},
          'cpu_baseline': {k: cb[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
          'e2e': {'value': cb['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
  print(json.dumps(line))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=200)
  ap.add_argument('--warmup', type=int, default=5)
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--workload', default='lr_giga_N1e7_S512', choices=sorted(WORKLOADS))
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--no-e2e', action='store_true')
  ap.add_argument('--also', default='lr_giga_N1e6_S256', help='second workload reported under "also" at 1 GPU')
  args = ap.parse_args()
  args.warmup = max(args.warmup, 3)
  if args.impl == 'reference':
    return run_reference(args)

  rank = int(os.environ.get('RANK', '0'))
  world = int(os.environ.get('WORLD_SIZE', '1'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  import bayesiancoresets_b200 as bc
  comm = None
  if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    comm = bc.comm.TorchComm()
  ctx = bc.Context.default(local_rank)

  def barrier():
    if comm is not None:
      comm.barrier()

  def max_over_ranks(x):
    if comm is None:
      return float(x)
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(x)], dtype=torch.float64, device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

  def measure(workload, steps, warmup, with_e2e):
    N, d, S = WORKLOADS[workload]
    lo, hi = bc.comm.even_shard(N, rank, world)
    Z, th_true = lr_shard(0, lo, hi, d)
    theta = lr_samples(0, th_true, S)
    prj = bc.LogisticRegressionProjector(lambda n, w, p: theta, S, ctx=ctx)
    kw = {'comm': comm} if comm is not None else {}
    cs = bc.HilbertCoreset(Z, prj, snnls=bc.snnls.GIGA, **kw)
    nat = cs.snnls._native
    cs.snnls.build(warmup)                                  # untimed warm-up iterations
    ctx.synchronize()
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
      clocks.start()
    cs.snnls.build(steps)                                   # timed: K iterations, one device-side loop
    ctx.synchronize()
    barrier()
    tm = nat.timing()
    clk = clocks.stop() if rank == 0 else None
    ok_steps = sum(1 for e in cs.snnls.last_events if e.code == 0)
    n_events = len(cs.snnls.last_events)
    build_ms = max_over_ranks(tm['build_ms'])
    if tm['scan_launches'] == 0:
      # persistent engine: the whole timed region is ONE launch of greedy_loop_kernel, which streams the
      # matrix `steps` times; its duration is the CUDA-event time of that launch on the library's stream
      kernel, launches_per_step, kernel_ms = 'greedy_loop_kernel', 1.0 / steps, build_ms
      bytes_per_launch = 4.0 * (hi - lo) * S * steps
    else:
      # launch-per-iteration engine: same loop continued with CUDA events around every scan launch
      # (kept out of the timed region above so the event records do not perturb it)
      nat.set_profiling(True)
      cs.snnls.build(min(steps, 50))
      ctx.synchronize()
      barrier()
      tp = nat.timing()
      nat.set_profiling(False)
      kernel, launches_per_step = 'scan_kernel', 1.0
      kernel_ms = max_over_ranks(tp['scan_ms'] / max(tp['scan_launches'], 1))
      bytes_per_launch = 4.0 * (hi - lo) * S
    res = {'N': N, 'd': d, 'S': S, 'build_ms': build_ms, 'kernel': kernel, 'kernel_ms': kernel_ms,
           'bytes_per_launch': bytes_per_launch, 'launches_per_step': launches_per_step, 'ok_steps': ok_steps,
           'events': n_events, 'rows_local': hi - lo, 'clocks': clk,
           'launches': tm['scan_launches'] + tm['step_launches'], 'error': cs.error(), 'size': int(cs.snnls.size())}
    if with_e2e:
      # end to end through the public API from HOST buffers: upload Z and theta, project on the device,
      # build(steps), read the coreset back -- everything inside the timed region
      del cs, nat
      Zp = bc.pinned_copy(Z)                                # the caller's input array, in page-locked host memory
      barrier()
      t0 = time.perf_counter()
      cs2 = bc.HilbertCoreset(Zp, prj, snnls=bc.snnls.GIGA, **kw)
      cs2.build(steps)
      wts, pts, idcs = cs2.get()
      err = cs2.error()
      ctx.synchronize()
      barrier()
      dt = max_over_ranks(time.perf_counter() - t0)
      res['e2e'] = {'value': steps / dt, 'unit': UNIT,
                    'h2d_bytes_per_step': int((Z.nbytes + theta.nbytes) / steps),
                    'd2h_bytes_per_step': int((wts.nbytes + idcs.nbytes + 48 * steps + 8) / steps),
                    'job': 'HilbertCoreset(Z_host [pinned], LR projector) + build(%d) + get() + error(): %.1f ms wall, '
                           'H2D %d bytes and D2H per job, amortised per step' % (steps, dt * 1e3, Z.nbytes + theta.nbytes)}
    return res

  r = measure(args.workload, args.steps, args.warmup, not args.no_e2e)
  peak, peak_src = measured_peak()
  achieved = r['bytes_per_launch'] / (r['kernel_ms'] * 1e-3) / 1e9
  ratio = measured_traffic_ratio(r['kernel'], r['S'])
  value = args.steps / (r['build_ms'] * 1e-3)
  line = {
    'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
    'ms_per_step': r['build_ms'] / args.steps, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
    'dtype': 'f32', 'data': 'synthetic',
    'config': {'workload': args.workload, 'N': r['N'], 'd': r['d'], 'S': r['S'], 'alg': 'GIGA',
               'sharding': 'N axis over %d GPU(s), %d rows/GPU' % (world, r['rows_local']),
               'l2': 'inputs larger than L2 (%.2f GB scanned per GPU per step)' % (4e-9 * r['rows_local'] * r['S']),
               'ok_steps': r['ok_steps'], 'final_error': r['error'], 'coreset_size': r['size']},
    'gpu_launches': r['launches'],
    'clocks': r['clocks'],
    'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                 'traffic': None if ratio is None else ratio * r['bytes_per_launch'],
                 'traffic_source': None if ratio is None else
                 'ncu dram__bytes_read+write per algorithmic byte (profiles/r01_loop_traffic.json) x bytes_per_launch',
                 'kernel': r['kernel'], 'bytes_per_launch': r['bytes_per_launch'],
                 'avg_launch_ms': r['kernel_ms'], 'peak_source': peak_src,
                 'kernel_share_of_step': r['kernel_ms'] * r['launches_per_step'] * args.steps / r['build_ms']},
  }
  if 'e2e' in r:
    line['e2e'] = r['e2e']
  if world == 1 and args.also in WORKLOADS and args.also != args.workload:
    a = measure(args.also, args.steps, args.warmup, False)
    agbs = a['bytes_per_launch'] / (a['kernel_ms'] * 1e-3) / 1e9
    line['also'] = {args.also: {'value': args.steps / (a['build_ms'] * 1e-3), 'unit': UNIT,
                                'ms_per_step': a['build_ms'] / args.steps, 'kernel': a['kernel'],
                                'kernel_gbs': agbs, 'roofline_frac': agbs / peak}}
  if rank == 0 and world == 1 and not args.no_cpu_baseline:
    cb, _ = cpu_reference_run(args.workload, 12, 2)
    line['cpu_baseline'] = {k: cb[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
  if rank == 0:
    print(json.dumps(line))
  if comm is not None:
    import torch.distributed as dist
    barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
  main()

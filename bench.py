#!/usr/bin/env python
"""Benchmark of the greedy sparse-NNLS hot path (BASELINE.json metric: greedy iters/sec on an
N x S projection).

  python bench.py --gpus 1 --steps K --warmup W            # this engine (CUDA, sm_100a)
  python bench.py --impl reference --steps K --warmup W    # the UNMODIFIED reference (baseline/_ref) on the host cores
  torchrun ... bench.py --gpus N ...                        # N-sharded over N GPUs of one node

Workload (config.workload): the north-star target -- synthetic logistic regression, N = 1e7
datapoints, d = 10, S = 512 posterior samples, GIGA.  A "step" is ONE greedy iteration (one full
pass of the scan kernel over the resident N x S matrix plus the reweight).  With --gpus N the
same N rows are sharded over the ranks (strong scaling): each rank scans N/ranks rows and the
candidates are exchanged over NVLink peer memory inside the loop kernel.

The other BASELINE.json configs ride along under "also" (each with its own roofline and CPU baseline):
  c2  LR HilbertCoreset GIGA N=1e6 S=256                 (configs[1])
  c3  Gaussian SparseVICoreset N=1e6 d=200 S=512         (configs[2]; 1 GPU)
  c4  LR HilbertCoreset OrthoPursuit N=1e7 S=512         (configs[3]; N-sharded with --gpus N)
  c5  Poisson BatchPSVICoreset gradient N=1e7 d=128      (configs[4]; N-sharded with --gpus N, all-reduce included)
Prints ONE JSON line on rank 0.
"""
import argparse
import hashlib
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
F64_NOMINAL_TFLOPS = 40.0      # B200 FP64 (tensor) peak, NVIDIA data sheet; used when no dgemm measurement is possible


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


def poisson_shard(lo, hi, d, seed=1):
  """rows [lo, hi) of the synthetic Poisson-regression data (generalising model_poiss.py:19-23): x = [N(0, I), 1],
  y ~ Poisson(softplus(x.th*)), z = [x, y]; block-seeded like lr_shard"""
  th_true = np.random.RandomState(0).randn(d) / np.sqrt(d)
  Z = np.empty((hi - lo, d + 1))
  blk = 500_000
  for b0 in range((lo // blk) * blk, hi, blk):
    r = np.random.RandomState(seed * 104729 + b0 // blk)
    X = np.hstack((r.randn(blk, d - 1), np.ones((blk, 1))))
    y = r.poisson(np.log1p(np.exp(X.dot(th_true)))).astype(np.float64)
    a, b = max(lo, b0), min(hi, b0 + blk)
    Z[a - lo:b - lo, :d] = X[a - b0:b - b0]
    Z[a - lo:b - lo, d] = y[a - b0:b - b0]
  return Z, th_true


def sel_hash(events):
  """order-sensitive hash of the selected GLOBAL indices (identical across 1 / 2 / 4 / 8 GPUs when the N-sharded runs
  select the same rows)"""
  f = np.array([e.f for e in events], dtype=np.int64)
  return hashlib.sha1(f.tobytes()).hexdigest()[:16]


class ClockSampler(object):
  """SM clock and throttle reasons sampled DURING the timed region: NVML in a thread (pynvml), nvidia-smi as a fallback.
  stop() returns only when the sampler is really gone -- a lingering nvidia-smi holds driver locks and was measured to stall
  the first jobs after the timed region by 0.3 - 1.3 s at 4 / 8 GPUs."""
  Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
       'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
       'clocks_event_reasons.sw_power_cap')
  BITS = {'hw_slowdown': 0x8, 'sw_power_cap': 0x4, 'sw_thermal_slowdown': 0x20, 'hw_thermal_slowdown': 0x40}

  def __init__(self, device):
    self.rows, self.proc, self.device = [], None, device
    self.sm, self.smax, self.reasons, self.stop_flag, self.t, self.nvml = [], [], set(), False, None, None

  def _nvml_loop(self):
    nv, h = self.nvml
    while not self.stop_flag:
      try:
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
        self.smax.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
        r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
        for nm, bit in self.BITS.items():
          if r & bit:
            self.reasons.add(nm)
      except Exception:
        pass
      time.sleep(0.005)

  def start(self):
    if os.environ.get('BCG_BENCH_NO_CLOCKS'):
      self.proc = None
      return
    try:
      import pynvml as nv
      nv.nvmlInit()
      vis = os.environ.get('CUDA_VISIBLE_DEVICES')
      idx = int(vis.split(',')[self.device]) if vis and vis.split(',')[0].isdigit() else self.device
      self.nvml = (nv, nv.nvmlDeviceGetHandleByIndex(idx))
      self.t = threading.Thread(target=self._nvml_loop, daemon=True)
      self.t.start()
      return
    except Exception:
      self.nvml = None
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
    if self.nvml is not None:
      self.stop_flag = True
      self.t.join(timeout=2)
      return {'sm_mhz': float(np.median(self.sm)) if self.sm else None, 'sm_max_mhz': max(self.smax) if self.smax else None,
              'reasons': sorted(self.reasons), 'samples': len(self.sm), 'source': 'nvml'}
    if self.proc is None:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
    time.sleep(0.15)
    self.proc.terminate()
    try:
      self.proc.wait(timeout=5)
    except Exception:
      self.proc.kill()
      self.proc.wait()
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
            'reasons': sorted(reasons), 'samples': len(sm), 'source': 'nvidia-smi'}


def measured_peak():
  try:
    with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
      return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
  except Exception:
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def measured_f64_peak(device):
  """float64 GEMM peak measured the way MEASURED_PEAKS.json measures bf16: cuBLAS dgemm (torch.matmul, 4096^3, best of
  5, CUDA events).  Falls back to the nominal data-sheet figure when torch cannot run it."""
  try:
    import torch
    with torch.cuda.device(device):
      a = torch.randn(4096, 4096, dtype=torch.float64, device='cuda')
      b = torch.randn(4096, 4096, dtype=torch.float64, device='cuda')
      torch.matmul(a, b)
      best = 1e9
      for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
      del a, b
      torch.cuda.empty_cache()
    return 2. * 4096**3 / (best * 1e-3) / 1e12, 'measured (cuBLAS dgemm 4096^3 via torch.matmul in this run, best of 5)'
  except Exception:
    return F64_NOMINAL_TFLOPS, 'nominal (B200 data sheet FP64 40 TFLOP/s; dgemm measurement unavailable)'


def measured_traffic_ratio(kernel, S, filter16=False):
  """DRAM bytes / algorithmic bytes of the dominant kernel from the committed `ncu --set full` capture
  (profiles/r02_loop_traffic.json, written by profiles/extract_traffic.py from tools/ncu_traffic_case.py; the round-1
  capture as a fallback); None when there is no capture of this kernel at this S"""
  try:
    with open(os.path.join(ROOT, 'profiles', 'r02_loop_traffic.json')) as f:
      for c in json.load(f)['captures']:
        if c['kernel'] == kernel and int(c['S']) == int(S) and bool(c.get('filter16', False)) == bool(filter16):
          return float(c['dram_bytes'])/float(c['algorithmic_bytes']), 'r02_loop_traffic.json'
  except Exception:
    pass
  try:
    with open(os.path.join(ROOT, 'profiles', 'r01_loop_traffic.json')) as f:
      t = json.load(f)
    if t.get('kernel') == kernel and int(t.get('S', -1)) == int(S) and not filter16:
      return float(t['dram_bytes'])/float(t['algorithmic_bytes']), 'r01_loop_traffic.json'
  except Exception:
    pass
  return None, None


# ------------------------------------------------------------------------------------------------------------------
# CPU side: the unmodified reference (baseline/_ref, else /root/reference), else the oracle port
# ------------------------------------------------------------------------------------------------------------------
def host_threads():
  cores = os.cpu_count()
  try:
    # torchrun exports OMP_NUM_THREADS=1: give the CPU reference every host core it can use
    import threadpoolctl
    threadpoolctl.threadpool_limits(limits=cores)
    used = [p.get('num_threads') for p in threadpoolctl.threadpool_info() if p.get('user_api') == 'blas']
    cores = max(used) if used else cores
  except Exception:
    pass
  return cores


def import_reference():
  """the unmodified reference package: baseline/_ref (pip --target install, travels to the GPU box), else /root/reference"""
  for path in (os.path.join(ROOT, 'baseline', '_ref'), '/root/reference'):
    if os.path.isdir(os.path.join(path, 'bayesiancoresets')):
      sys.path.insert(0, path)
      try:
        import bayesiancoresets as ref
        return ref, path
      except Exception:
        sys.path.remove(path)
  return None, None


def cpu_greedy_run(N, d, S, alg, steps, warmup, sample_rows):
  """CPU arm of the greedy loop on a bounded row sample of the same workload: HilbertCoreset(Z, BlackBoxProjector(...,
  model_lr.log_likelihood), snnls=alg) of the UNMODIFIED reference when it is importable (kind 'reference'), else the
  oracle port (kind 'port', bit-identical to the reference).  The loop is memory-bound and linear in N, so iters/s is
  scaled by sample/N (N = 1e7, S = 512 needs 82 GB in float64)."""
  from oracle import greedy, models
  cores = host_threads()
  Z, th_true = lr_shard(0, 0, sample_rows, d)
  theta = lr_samples(0, th_true, S)
  ref, where = import_reference()
  t0 = time.perf_counter()
  if ref is not None:
    kind = 'reference'
    prj = ref.BlackBoxProjector(lambda n, w, p: theta, S, models.lr_loglik)    # model_lr.log_likelihood is user code
    cls = {'GIGA': ref.snnls.GIGA, 'FrankWolfe': ref.snnls.FrankWolfe, 'OrthoPursuit': ref.snnls.OrthoPursuit}[alg]
    cs = ref.HilbertCoreset(Z, prj, snnls=cls)
    build = cs.build
  else:
    kind = 'port'
    vecs = models.project(models.lr_loglik, Z, theta)
    o = {'GIGA': greedy.GigaOracle, 'FrankWolfe': greedy.FrankWolfeOracle, 'OrthoPursuit': greedy.OrthoPursuitOracle}[alg](
        vecs.T, vecs.sum(axis=0))
    build = o.build
  t_setup = time.perf_counter() - t0
  build(max(warmup, 1))
  t0 = time.perf_counter()
  build(steps)
  dt = time.perf_counter() - t0
  its = steps / dt
  return {'value': its * sample_rows / N, 'unit': UNIT, 'cores': cores, 'kind': kind,
          'sample': '%d of %d rows (S=%d, float64), %s %s: %d timed iterations at %.3f s/iter on the sample after %d '
                    'warm-up; iters/s scaled by sample/N; projection + solver construction on the sample took %.1f s (not '
                    'in the metric)' % (sample_rows, N, S, 'unmodified reference (' + str(where) + ')' if ref is not None
                                        else 'oracle port', alg, steps, dt / steps, max(warmup, 1), t_setup),
          'sample_iters_per_s': its}


def run_reference(args):
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  steps = min(args.steps, 40)
  warmup = min(args.warmup, 5)
  N, d, S = WORKLOADS[args.workload]
  cb = cpu_greedy_run(N, d, S, 'GIGA', steps, warmup, min(N, args.ref_rows))
  line = {'impl': 'reference', 'metric': METRIC, 'value': cb['value'], 'unit': UNIT, 'n_gpus': args.gpus,
          'steps': steps, 'warmup': warmup, 'ms_per_step': 1e3 / cb['value'],
          'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
          'config': {'workload': args.workload, 'N': N, 'd': d, 'S': S, 'alg': 'GIGA'},
          'cpu_baseline': {k: cb[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
          'e2e': {'value': cb['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
  print(json.dumps(line))


def cpu_projection_pass(model, N, d, S, sample_rows):
  """CPU arm of one `project(data).sum(axis=0)` pass (sparsevi.py:71-72 / bpsvi.py:49-51) on a row sample: the reference's
  BlackBoxProjector.project with the oracle's float64 model callback; seconds scaled by N/sample."""
  from oracle import models
  cores = host_threads()
  rng = np.random.RandomState(0)
  ref, where = import_reference()
  if model == 'gaussian':
    x = rng.randn(sample_rows, d) + 1.
    theta = rng.randn(S, d)
    Si = np.eye(d)
    f = lambda a, t: models.gaussian_loglik(a, t, Si, 0.)
  else:
    x, th_true = poisson_shard(0, sample_rows, d)
    theta = th_true + 0.05 * np.random.RandomState(2).randn(S, d)
    f = models.poisson_loglik
  if ref is not None:
    prj = ref.BlackBoxProjector(lambda n, w, p: theta, S, f)
    run = lambda: prj.project(x).sum(axis=0)
  else:
    run = lambda: models.project(f, x, theta).sum(axis=0)
  run()
  t0 = time.perf_counter()
  reps = 3
  for _ in range(reps):
    run()
  dt = (time.perf_counter() - t0) / reps
  return {'value': dt * N / sample_rows, 'unit': 's/pass', 'cores': cores, 'kind': 'reference' if ref is not None else 'port',
          'sample': '%d of %d rows (d=%d, S=%d, float64 %s log-likelihood through %s): %.3f s per project().sum(0) pass on the '
                    'sample, scaled by N/sample' % (sample_rows, N, d, S, model, 'the reference BlackBoxProjector' if ref is not
                                                     None else 'the oracle port', dt)}


# ------------------------------------------------------------------------------------------------------------------
def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=200)
  ap.add_argument('--warmup', type=int, default=5)
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--workload', default='lr_giga_N1e7_S512', choices=sorted(WORKLOADS))
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--no-e2e', action='store_true')
  ap.add_argument('--also', default='c2,c3,c4,c5', help='comma list of the other BASELINE configs to report under "also" (or "none")')
  ap.add_argument('--ref-rows', type=int, default=1_000_000, help='rows of the CPU reference sample (BASELINE.md section 4)')
  ap.add_argument('--scale', type=float, default=1.0, help='shrink the "also" configs (functional runs)')
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
    comm = bc.comm.default_comm()
  ctx = bc.Context.default(local_rank)
  also = [] if args.also in ('', 'none') else [a.strip() for a in args.also.split(',')]

  def barrier():
    if comm is not None:
      comm.barrier()

  def max_over_ranks(x):
    return float(x) if comm is None else float(comm.allreduce_max(np.array([float(x)]))[0])

  kw = {'comm': comm} if comm is not None else {}
  peak, peak_src = measured_peak()

  def greedy_measure(name, N, d, S, alg, steps, warmup, with_e2e, keep=None):
    """one HilbertCoreset greedy workload: returns the result dict; `keep` = (Z, theta, prj, vecs-holding coreset) of a
    previous call on the same data, so that c4 (OMP on the headline matrix) does not regenerate it"""
    lo, hi = bc.comm.even_shard(N, rank, world)
    if keep is None:
      Z, th_true = lr_shard(0, lo, hi, d)
      theta = lr_samples(0, th_true, S)
      prj = bc.LogisticRegressionProjector(lambda n, w, p: theta, S, ctx=ctx)
    else:
      Z, theta, prj = keep
    cls = {'GIGA': bc.snnls.GIGA, 'FrankWolfe': bc.snnls.FrankWolfe, 'OrthoPursuit': bc.snnls.OrthoPursuit}[alg]
    cs = bc.HilbertCoreset(Z, prj, snnls=cls, **kw)
    nat = cs.snnls._native
    cs.snnls.build(warmup)                                  # untimed warm-up iterations
    ev_all = list(cs.snnls.last_events)
    ctx.synchronize()
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
      clocks.start()
    f16_on, f16_rows0 = nat.filter16_stats()
    cs.snnls.build(steps)                                   # timed: K iterations, one device-side loop
    ctx.synchronize()
    barrier()
    tm = nat.timing()
    f16_rows = nat.filter16_stats()[1] - f16_rows0          # rows re-scanned in float32 inside the timed region (this rank)
    clk = clocks.stop() if rank == 0 else None
    ev_all += list(cs.snnls.last_events)
    ok_steps = sum(1 for e in cs.snnls.last_events if e.code == 0)
    build_ms = max_over_ranks(tm['build_ms'])
    if tm['scan_launches'] == 0:
      # persistent engine: the whole timed region is ONE launch of greedy_loop_kernel / omp_loop_kernel, which streams the
      # matrix `steps` times; its duration is the CUDA-event time of that launch on the library's stream
      kernel, launches_per_step, kernel_ms = ('omp_loop_kernel' if alg == 'OrthoPursuit' else 'greedy_loop_kernel'), 1.0 / steps, build_ms
      bytes_per_launch = 4.0 * (hi - lo) * S * steps
    else:
      # launch-per-iteration engine (OMP; GIGA / FW after an exact-selection stop): same loop continued with CUDA
      # events around every scan launch (kept out of the timed region above so the records do not perturb it)
      nat.set_profiling(True)
      cs.snnls.build(min(steps, 50))
      ctx.synchronize()
      barrier()
      tp = nat.timing()
      nat.set_profiling(False)
      kernel, launches_per_step = 'scan_kernel', 1.0
      kernel_ms = max_over_ranks(tp['scan_ms'] / max(tp['scan_launches'], 1))
      bytes_per_launch = 4.0 * (hi - lo) * S
    achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
    f16 = bool(f16_on) and tm['scan_launches'] == 0
    ratio, ratio_src = measured_traffic_ratio(kernel, S, f16)
    res = {'N': N, 'd': d, 'S': S, 'alg': alg, 'build_ms': build_ms, 'rows_local': hi - lo, 'clocks': clk,
           'ok_steps': ok_steps, 'launches': tm['scan_launches'] + tm['step_launches'], 'error': cs.error(),
           'size': int(cs.snnls.size()), 'sel_hash': sel_hash(ev_all), 'exact_selections': nat.exact_count(), 'filter16': bool(f16),
           'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                        'traffic': None if ratio is None else ratio * bytes_per_launch,
                        'traffic_source': None if ratio is None else
                        'ncu dram__bytes_read+write per algorithmic byte (profiles/%s) x bytes_per_launch' % ratio_src,
                        'kernel': kernel, 'bytes_per_launch': bytes_per_launch, 'avg_launch_ms': kernel_ms,
                        'peak_source': peak_src,
                        'kernel_share_of_step': kernel_ms * launches_per_step * steps / build_ms},
           'keep': (Z, theta, prj)}
    if f16:
      # float16 pre-filter (csrc/filter_bounds.h): `achieved` keeps SURVEY 8(d)'s algorithmic figure (4 N S bytes per iteration,
      # one float32 read of the matrix), which this kernel no longer has to move -- it streams a float16 copy (2 bytes per
      # element) and re-reads in float32 only the row groups whose bound reaches the maximum.  `moved` is the roofline of the
      # bytes it does move.
      ld16 = (S + 7)//8*8
      moved = 2.0*(hi - lo)*ld16*steps + 4.0*S*f16_rows
      res['roofline']['filter16'] = {
        'moved_bytes_per_launch': moved, 'achieved': moved/(kernel_ms*1e-3)/1e9, 'unit': 'GB/s',
        'frac': moved/(kernel_ms*1e-3)/1e9/peak, 'rows_rescanned_float32_per_step': f16_rows/float(steps),
        'note': 'selection bit-identical to the float32 stream (config.sel_hash; tests/test_gpu_parity.py::test_filter16_*); '
                'roofline.achieved / frac above use the 4 N S algorithmic bytes of SURVEY 8(d) and therefore exceed the HBM peak'}
    if with_e2e:
      # end to end through the public API from HOST buffers: upload Z and theta, project on the device,
      # build(steps), read the coreset back -- everything inside the timed region.  Timed twice: from a page-locked
      # source (bc.pinned_copy: DMA straight from the caller's array) and from an ordinary pageable ndarray (what a
      # reference user holds; staged through the library's pinned buffers)
      del cs, nat
      out, runs = {}, {}
      reps = 5
      for label, src in (('pinned', bc.pinned_copy(Z)), ('pageable', Z)):
        runs[label] = []
        for _ in range(reps):                               # the job is run 5 times; the BEST is reported, all 5 are listed
          # in `job` (wall clock on a shared host: stalls of 0.1 - 1 s in single jobs were observed on some boxes, with
          # every phase of the job clean when re-timed -- tools/e2e_diag.py; MEASURED_PEAKS.json is best-of-10 as well)
          barrier()
          t0 = time.perf_counter()
          cs2 = bc.HilbertCoreset(src, prj, snnls=cls, **kw)
          cs2.build(steps)
          wts, pts, idcs = cs2.get()
          err = cs2.error()
          ctx.synchronize()
          barrier()
          runs[label].append(max_over_ranks(time.perf_counter() - t0))
          d2h = int(wts.nbytes + idcs.nbytes + 48 * steps + 8)
          del cs2
        out[label] = float(np.min(runs[label]))
        del src
      fmt = lambda xs: ' / '.join('%.1f' % (x * 1e3) for x in xs)
      res['e2e'] = {'value': steps / out['pinned'], 'unit': UNIT,
                    'h2d_bytes_per_step': int((Z.nbytes + theta.nbytes) / steps),
                    'd2h_bytes_per_step': int(d2h / steps),
                    'pageable_value': steps / out['pageable'],
                    'job': 'HilbertCoreset(Z_host, LR projector) + build(%d) + get() + error(), wall clock, max over ranks, best of '
                           '%d runs: %s ms from a page-locked source (value), %s ms from a pageable ndarray (pageable_value); H2D %d '
                           'bytes and D2H per job, amortised per step' % (steps, reps, fmt(runs['pinned']), fmt(runs['pageable']),
                                                                         Z.nbytes + theta.nbytes)}
    return res

  N, d, S = WORKLOADS[args.workload]
  r = greedy_measure(args.workload, N, d, S, 'GIGA', args.steps, args.warmup, not args.no_e2e)
  line = {
    'metric': METRIC, 'value': args.steps / (r['build_ms'] * 1e-3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
    'warmup': args.warmup, 'ms_per_step': r['build_ms'] / args.steps, 'higher_is_better': True, 'scaling': 'strong',
    'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
    'config': {'workload': args.workload, 'N': N, 'd': d, 'S': S, 'alg': 'GIGA',
               'sharding': 'N axis over %d GPU(s), %d rows/GPU' % (world, r['rows_local']),
               'l2': ('inputs larger than L2 (%.2f GB float16 copy of the %.2f GB matrix scanned per GPU per step)' %
                      (2e-9*r['rows_local']*((S + 7)//8*8), 4e-9*r['rows_local']*S)) if r['filter16'] else
                     'inputs larger than L2 (%.2f GB scanned per GPU per step)' % (4e-9 * r['rows_local'] * S),
               'ok_steps': r['ok_steps'], 'final_error': r['error'], 'coreset_size': r['size'], 'sel_hash': r['sel_hash'],
               'exact_selections': r['exact_selections'], 'filter16': r['filter16']},
    'gpu_launches': r['launches'], 'clocks': r['clocks'], 'roofline': r['roofline'],
  }
  if 'e2e' in r:
    line['e2e'] = r['e2e']
  want_cpu = rank == 0 and world == 1 and not args.no_cpu_baseline
  extra = {}

  # ---- c4: OrthoPursuit on the headline matrix (BASELINE configs[3]; N-sharded under torchrun) -----------------
  if 'c4' in also and args.workload == 'lr_giga_N1e7_S512':
    a = greedy_measure('c4', N, d, S, 'OrthoPursuit', args.steps, args.warmup, False, keep=r['keep'])
    extra['c4_lr_omp_N1e7_S512'] = {
      'metric': METRIC, 'value': args.steps / (a['build_ms'] * 1e-3), 'unit': UNIT, 'ms_per_step': a['build_ms'] / args.steps,
      'n_gpus': world, 'config': {'alg': 'OrthoPursuit', 'N': N, 'S': S, 'rows_local': a['rows_local'], 'sel_hash': a['sel_hash'],
                                  'final_error': a['error'], 'coreset_size': a['size'], 'exact_selections': a['exact_selections']},
      'roofline': a['roofline'], 'gpu_launches': a['launches']}
  del r

  # ---- c2: BASELINE configs[1] ---------------------------------------------------------------------------------
  if 'c2' in also and world == 1:
    n2 = int(1_000_000 * args.scale)
    a = greedy_measure('c2', n2, 10, 256, 'GIGA', args.steps, args.warmup, False)
    extra['c2_lr_giga_N1e6_S256'] = {
      'metric': METRIC, 'value': args.steps / (a['build_ms'] * 1e-3), 'unit': UNIT, 'ms_per_step': a['build_ms'] / args.steps,
      'config': {'alg': 'GIGA', 'N': n2, 'S': 256, 'sel_hash': a['sel_hash'], 'final_error': a['error']},
      'roofline': a['roofline']}
    del a

  f64_peak = None
  if ('c3' in also and world == 1) or 'c5' in also:
    f64_peak = measured_f64_peak(local_rank)

  def k3b_roofline(n_rows, dd, din, SS, sec, kernel):
    """column-sum-only projection: flops 2 N d S against the float64 GEMM peak; bytes 8 N d_in against HBM"""
    tf = 2. * n_rows * dd * SS / sec / 1e12
    gbs = 8. * n_rows * din / sec / 1e9
    return {'bound': 'tensor', 'achieved': tf, 'peak': f64_peak[0], 'unit': 'TFLOP/s', 'frac': tf / f64_peak[0],
            'traffic': None, 'kernel': kernel, 'flops_per_launch': 2. * n_rows * dd * SS, 'avg_launch_ms': sec * 1e3,
            'peak_source': f64_peak[1],
            'hbm': {'achieved': gbs, 'peak': peak, 'unit': 'GB/s', 'frac': gbs / peak, 'bytes_per_launch': 8. * n_rows * din}}

  # ---- c3: Gaussian SparseVI (BASELINE configs[2]) -------------------------------------------------------------
  if 'c3' in also and world == 1:
    n3, d3, S3, opt_itrs = int(1_000_000 * args.scale), 200, 512, 100
    rng = np.random.RandomState(0)
    x = rng.randn(n3, d3) + 1.                                   # examples/gaussian/main.py:72,82: N(1_d, I)
    th0, Sig0inv, Siginv = np.zeros(d3), np.eye(d3), np.eye(d3)

    # the sampler_w of examples/gaussian/main.py:107-113 (weighted_post, model_gaussian.py:23-30): normals drawn on the
    # host in the reference's order, factorisation and sample transform on the device (csrc/sampler_kernels.cuh)
    sampler_w = bc.GaussianPosteriorSampler(th0, Sig0inv, Siginv, ctx=ctx)
    np.random.seed(0)
    prj = bc.GaussianProjector(sampler_w, S3, Siginv, ctx=ctx)
    prj.project_sum(x)                                           # one-off upload of x + warm-up
    ctx.synchronize()
    ts = []
    for _ in range(10):
      t0 = time.perf_counter()
      prj.project_sum(x)
      ts.append(time.perf_counter() - t0)
    t_sum = float(np.median(ts))                                 # host-driven calls on a shared box: medians, not means
    v = prj.project_device(x, cache=True)                        # warm-up (first use of the streamed DMMA projection kernel)
    del v
    ctx.synchronize()
    t0 = time.perf_counter()
    v = prj.project_device(x, cache=True)
    ctx.synchronize()
    t_full = time.perf_counter() - t0
    del v
    svi = bc.SparseVICoreset(x, prj, opt_itrs=opt_itrs)
    svi.build(1)
    ts = []
    for _ in range(3):
      t0 = time.perf_counter()
      svi.build(1)
      ts.append(time.perf_counter() - t0)
    t_iter = float(np.median(ts))
    extra['c3_gaussian_sparsevi_N1e6_d200_S512'] = {
      'metric': 'sparsevi_build_iter_seconds', 'value': t_iter, 'unit': 's/iter', 'higher_is_better': False,
      'config': {'N': n3, 'd': d3, 'S': S3, 'opt_itrs': opt_itrs, 'coreset_size': int(svi.size()),
                 'colsum_pass_s': t_sum, 'materialising_pass_s': t_full, 'build_iter_s_runs': [round(t, 4) for t in ts],
                 'note': 'one build iteration = 1 materialising projection + correlation arg-max + opt_itrs column-sum passes '
                         '(sparsevi.py:16-76); whole C-ABI calls, wall clock'},
      'roofline': k3b_roofline(n3, d3, d3, S3, t_sum, 'project_sum_mma_kernel<LINEAR> (whole bcg_dataset_project call)')}
    if want_cpu:
      ns = max(2000, int(20_000 * min(1., args.scale * 10)))
      cp = cpu_projection_pass('gaussian', n3, d3, S3, ns)
      cp_iter = dict(cp)
      cp_iter['value'] = cp['value'] * (opt_itrs + 1)
      cp_iter['unit'] = 's/iter'
      cp_iter['sample'] += '; a build iteration is (1 + opt_itrs) = %d such passes (94 %% of its time, SURVEY 3.3)' % (opt_itrs + 1)
      extra['c3_gaussian_sparsevi_N1e6_d200_S512']['cpu_baseline'] = cp_iter
    del svi, prj, x

  # ---- c5: Poisson BatchPSVI gradient (BASELINE configs[4]; N-sharded under torchrun, all-reduce included) -------
  if 'c5' in also:
    n5, d5, S5, sz = int(10_000_000 * args.scale), 128, 512, 100
    lo, hi = bc.comm.even_shard(n5, rank, world)
    Z, th_true = poisson_shard(lo, hi, d5)
    theta = th_true + 0.05 * np.random.RandomState(2).randn(S5, d5)
    prj = bc.PoissonProjector(lambda n, w, p: theta, S5, ctx=ctx)
    bp = bc.BatchPSVICoreset(Z, prj, opt_itrs=1, **kw)
    x0 = np.hstack((np.full(sz, n5 / sz), poisson_shard(0, sz, d5)[0].reshape(-1)))
    bp.gradient(x0.copy(), sz, d5 + 1)                           # one-off upload of Z + warm-up
    ctx.synchronize()
    barrier()
    tg = []
    for _ in range(5):
      barrier()
      t0 = time.perf_counter()
      g = bp.gradient(x0.copy(), sz, d5 + 1)
      ctx.synchronize()
      barrier()
      tg.append(max_over_ranks(time.perf_counter() - t0))
    dt = float(np.median(tg))
    extra['c5_poisson_bpsvi_grad_N1e7_d128_S512'] = {
      'metric': 'bpsvi_gradient_seconds', 'value': dt, 'unit': 's/grad', 'higher_is_better': False, 'n_gpus': world,
      'config': {'N': n5, 'rows_local': hi - lo, 'd': d5, 'S': S5, 'K': sz, 'grad_norm': float(np.linalg.norm(g)), 's_per_grad_runs': [round(t, 4) for t in tg],
                 'note': 'one grd() evaluation of bpsvi.py:46-55: column-sum projection of the local shard, S-vector all-reduce '
                         'over the ranks, K pseudo-point projection + gradient contraction; wall clock, max over ranks'},
      'roofline': k3b_roofline(hi - lo, d5, d5 + 1, S5, dt, 'project_sum_mma_kernel<POISSON> (whole gradient evaluation)')}
    if want_cpu:
      ns = max(2000, int(20_000 * min(1., args.scale * 10)))
      cp = cpu_projection_pass('poisson', n5, d5, S5, ns)
      cp['unit'] = 's/grad'
      cp['sample'] += '; the gradient step is 94 % this pass (SURVEY 3.4)'
      extra['c5_poisson_bpsvi_grad_N1e7_d128_S512']['cpu_baseline'] = cp
    del bp, prj, Z

  if want_cpu:
    cb = cpu_greedy_run(N, d, S, 'GIGA', 12, 2, min(N, args.ref_rows))
    line['cpu_baseline'] = {k: cb[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
    if 'c4_lr_omp_N1e7_S512' in extra:
      co = cpu_greedy_run(N, d, S, 'OrthoPursuit', 12, 2, min(N, args.ref_rows // 2))
      extra['c4_lr_omp_N1e7_S512']['cpu_baseline'] = {k: co[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
    if 'c2_lr_giga_N1e6_S256' in extra:
      c2 = cpu_greedy_run(1_000_000, 10, 256, 'GIGA', 12, 2, min(1_000_000, args.ref_rows // 2))
      extra['c2_lr_giga_N1e6_S256']['cpu_baseline'] = {k: c2[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
  if extra:
    line['also'] = extra
  if rank == 0:
    print(json.dumps(line))
  if comm is not None:
    barrier()
    comm.close()


if __name__ == '__main__':
  main()

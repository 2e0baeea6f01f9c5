"""launches whose DRAM traffic bench.py's roofline.traffic refers to (run under `ncu --set full -k regex:greedy_loop_kernel|omp_loop_kernel|scan_kernel`):
GIGA loop kernel at N=1e7 S=512 (5 iterations in one launch), OMP scan kernel at the same size, GIGA loop kernel at N=1e6 S=256"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bayesian-coresets_b200')); sys.path.insert(0, ROOT)
import bayesiancoresets_b200 as bc
from bench import lr_shard, lr_samples
for N, S, runs in ((10_000_000, 512, (('GIGA', 5, '1'), ('OrthoPursuit', 2, '1'), ('OrthoPursuit', 2, '0'))), (1_000_000, 256, (('GIGA', 5, '1'),))):
  Z, th = lr_shard(0, 0, N, 10)
  theta = lr_samples(0, th, S)
  prj = bc.LogisticRegressionProjector(lambda n, w, p: theta, S)
  for alg, it, omp_loop in runs:
    os.environ['BCG_OMP_LOOP'] = omp_loop            # persistent OMP kernel, then the launch-per-iteration scan kernel
    cs = bc.HilbertCoreset(Z, prj, snnls=getattr(bc.snnls, alg))
    cs.build(it)
    print(alg, N, S, it, cs.error(), flush=True)
    del cs

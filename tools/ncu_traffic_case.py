"""launches whose DRAM traffic bench.py's roofline.traffic refers to (run under
`ncu --set full -k regex:greedy_loop_kernel|omp_loop_kernel|scan_kernel`), in this order:
  N=1e7 S=512: GIGA loop kernel with the float16 pre-filter (5 iterations in one launch), OMP loop kernel with it (2),
               GIGA loop kernel / OMP loop kernel streaming float32 (BCG_FILTER16=0), OMP launch-per-iteration scan kernel (2 launches)
  N=1e6 S=256: GIGA loop kernel with and without the pre-filter"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bayesian-coresets_b200')); sys.path.insert(0, ROOT)
import bayesiancoresets_b200 as bc
from bench import lr_shard, lr_samples
CASES = ((10_000_000, 512, (('GIGA', 5, '1', '1'), ('OrthoPursuit', 2, '1', '1'), ('GIGA', 5, '1', '0'), ('OrthoPursuit', 2, '1', '0'),
                            ('OrthoPursuit', 2, '0', '0'))),
         (1_000_000, 256, (('GIGA', 5, '1', '1'), ('GIGA', 5, '1', '0'))))
for N, S, runs in CASES:
  Z, th = lr_shard(0, 0, N, 10)
  theta = lr_samples(0, th, S)
  prj = bc.LogisticRegressionProjector(lambda n, w, p: theta, S)
  for alg, it, omp_loop, f16 in runs:
    os.environ['BCG_OMP_LOOP'] = omp_loop            # persistent OMP kernel, or the launch-per-iteration scan kernel
    os.environ['BCG_FILTER16'] = f16                 # read when the solver is created
    cs = bc.HilbertCoreset(Z, prj, snnls=getattr(bc.snnls, alg))
    cs.build(it)
    print(alg, N, S, it, f16, cs.error(), cs.snnls._native.filter16_stats(), flush=True)
    del cs

"""diagnostic: per-iteration device timeline of the persistent loop kernel"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bayesian-coresets_b200')); sys.path.insert(0, ROOT)
import numpy as np
import bayesiancoresets_b200 as bc
from bench import lr_shard, lr_samples, WORKLOADS
wl = sys.argv[1] if len(sys.argv) > 1 else 'lr_giga_N1e6_S256'
N, d, S = WORKLOADS[wl]
Z, th = lr_shard(0, 0, N, d)
theta = lr_samples(0, th, S)
comm = None
if int(os.environ.get('WORLD_SIZE', '1')) > 1:
  import torch, torch.distributed as dist
  lr_ = int(os.environ['LOCAL_RANK']); torch.cuda.set_device(lr_)
  dist.init_process_group('nccl', device_id=torch.device('cuda', lr_))
  comm = bc.comm.TorchComm()
  lo, hi = bc.comm.even_shard(N, comm.rank, comm.world)
  Z = Z[lo:hi]
prj = bc.LogisticRegressionProjector(lambda n, w, p: theta, S)
cs = bc.HilbertCoreset(Z, prj, **({'comm': comm} if comm else {}))
cs.snnls.build(5)
nat = cs.snnls._native
nat.set_trace(True)
cs.snnls.build(40)
t = nat.trace().astype(np.int64)
t0 = t[0, 2]
if comm is not None and comm.rank != 0:
  comm.barrier(); sys.exit(0)
print('build_ms', nat.timing()['build_ms'])
print('it  go_seen  scan_end(cta0w0)  arrived  published   | scan_us  wait_grid_us  control_us  go_latency_us')
for i in range(min(len(t), 12)):
  arrived, pub, go, send, c4, c5, c6, c7 = t[i]
  nxt_go = t[i+1, 2] if i + 1 < len(t) else pub
  print('%2d %9.1f %9.1f %9.1f %9.1f | %7.1f %7.1f %7.1f %7.1f' % (i, (go-t0)/1e3, (send-t0)/1e3, (arrived-t0)/1e3, (pub-t0)/1e3,
        (send-go)/1e3, (arrived-send)/1e3, (pub-arrived)/1e3, (nxt_go-pub)/1e3),
        '| cands %.1f row %.1f search %.1f apply %.1f dir+pub %.1f' % ((c4-arrived)/1e3, (c5-c4)/1e3, (c6-c5)/1e3, (c7-c6)/1e3, (pub-c7)/1e3))

if comm is not None:
  comm.barrier()

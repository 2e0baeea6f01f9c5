"""diagnostic: per-iteration device timeline of the persistent loop kernel, on 1 or N GPUs (one process per GPU; any
launcher that exports RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT; no torch).  Every rank prints the mean
duration of each segment of an iteration; rank 0 also prints the first iterations in full.
   python tools/trace_loop.py lr_giga_N1e7_S512 [iters]"""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bayesian-coresets_b200')); sys.path.insert(0, ROOT)
import numpy as np
import bayesiancoresets_b200 as bc
from bench import lr_shard, lr_samples, WORKLOADS
wl = sys.argv[1] if len(sys.argv) > 1 else 'lr_giga_N1e6_S256'
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 60
N, d, S = WORKLOADS[wl]
comm = bc.comm.default_comm()
lo, hi = bc.comm.even_shard(N, comm.rank, comm.world)
Z, th = lr_shard(0, lo, hi, d)
theta = lr_samples(0, th, S)
prj = bc.LogisticRegressionProjector(lambda n, w, p: theta, S)
cs = bc.HilbertCoreset(Z, prj, **({'comm': comm} if comm.world > 1 else {}))
cs.snnls.build(5)
nat = cs.snnls._native
nat.set_trace(True)
cs.snnls.build(iters)
t = nat.trace().astype(np.int64)
build_ms = nat.timing()['build_ms']
# trace columns: 0 grid arrived, 1 next direction published, 2 CTA0/warp0 saw go, 3 CTA0/warp0 scan end,
#                4 candidates reduced, 5 winning row fetched (N-sharded: after the mailbox exchange), 6 line search, 7 committed
seg = {'scan_us (go seen -> CTA 0 warp 0 done)': t[:, 3] - t[:, 2],
       'grid_tail_us (CTA 0 done -> all CTAs arrived)': t[:, 0] - t[:, 3],
       'candidates_us': t[:, 4] - t[:, 0],
       'row_fetch_and_exchange_us': t[:, 5] - t[:, 4],
       'line_search_us': t[:, 6] - t[:, 5],
       'apply_us': t[:, 7] - t[:, 6],
       'direction_and_publish_us': t[:-1, 1] - t[:-1, 7],
       'go_latency_us (published -> scan warps run)': t[1:, 2] - t[:-1, 1],
       'iteration_us (go seen -> next go seen)': t[1:, 2] - t[:-1, 2]}
out = {'rank': comm.rank, 'world': comm.world, 'workload': wl, 'rows_local': hi - lo, 'iters': iters, 'build_ms': build_ms,
       'ms_per_iter': build_ms/iters, 'ideal_scan_us_at_6543GBs': 4.*(hi - lo)*S/6543.1e9*1e6,
       'mean_us': {k: round(float(np.mean(v))/1e3, 2) for k, v in seg.items()},
       'p90_us': {k: round(float(np.percentile(v, 90))/1e3, 2) for k, v in seg.items()}}
for r in range(comm.world):
  comm.barrier()
  if r == comm.rank:
    print(json.dumps(out), flush=True)
comm.barrier()
if comm.rank == 0:
  t0 = t[0, 2]
  print('it  go_seen  scan_end(cta0w0)  arrived  published   | scan_us  wait_grid_us  control_us  go_latency_us')
  for i in range(min(len(t) - 1, 8)):
    arrived, pub, go, send, c4, c5, c6, c7 = t[i]
    print('%2d %9.1f %9.1f %9.1f %9.1f | %7.1f %7.1f %7.1f %7.1f' % (i, (go-t0)/1e3, (send-t0)/1e3, (arrived-t0)/1e3, (pub-t0)/1e3,
          (send-go)/1e3, (arrived-send)/1e3, (pub-arrived)/1e3, (t[i+1, 2]-pub)/1e3),
          '| cands %.1f row+exchange %.1f search %.1f apply %.1f dir+pub %.1f' % ((c4-arrived)/1e3, (c5-c4)/1e3, (c6-c5)/1e3, (c7-c6)/1e3, (pub-c7)/1e3), flush=True)
comm.barrier()
comm.close()

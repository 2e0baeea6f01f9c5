"""per-CUDA-source-line stall samples from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name ...`:
   python tools/ncu_lines.py src.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
cur, hdr, out = None, None, []
STALLS = ['stall_wait', 'stall_long_sb', 'stall_math', 'stall_short_sb', 'stall_barrier', 'stall_mio', 'stall_not_selected',
          'stall_selected', 'stall_lg', 'stall_dispatch', 'stall_branch_resolving', 'stall_no_inst']
def num(x):
  try: return float(x)
  except Exception: return 0.
for r in rows:
  if not r: continue
  if r[0] in ('File Path', 'File Name'): cur = r[1].split('/')[-1]; continue
  if r[0] == 'Line No': hdr = r; continue
  if hdr is None or r[0] == '': continue
  try: ln = int(r[0])
  except Exception: continue
  d = {h: r[i] for i, h in enumerate(hdr) if i < len(r)}
  out.append((num(d.get('# Samples')), num(d.get('Instructions Executed')), cur, ln, r[1].strip()[:90], {s: num(d.get(s)) for s in STALLS}))
ts, ti = sum(o[0] for o in out), sum(o[1] for o in out)
tot = {s: sum(o[5][s] for o in out) for s in STALLS}
print('samples %d, warp instructions %d' % (ts, ti))
print('stall mix: ' + ', '.join('%s %.1f%%' % (s[6:], 100*tot[s]/max(ts, 1)) for s in STALLS if tot[s] > 0.005*ts))
for o in sorted(out, key=lambda o: -o[0])[:top]:
  mix = ' '.join('%s=%d' % (s[6:], o[5][s]) for s in STALLS if o[5][s] > 0.1*o[0])
  print('%5.1f%% smp %5.1f%% ins  %s:%d  %s   [%s]' % (100*o[0]/max(ts, 1), 100*o[1]/max(ti, 1), o[2], o[3], o[4], mix))

"""per-CUDA-source-line instruction counts and stall samples from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur, out, hdr = None, [], None
def num(x):
  try: return int(float(x))
  except Exception: return 0
for r in rows:
  if not r: continue
  if r[0] in ('File Path', 'File Name'): cur = r[1]; continue
  if r[0] == 'Line No': hdr = r; continue
  if hdr is None or len(r) < 8 or r[0] == '': continue
  try: ln = int(r[0])
  except Exception: continue
  ie, sm = hdr.index('Instructions Executed'), hdr.index('# Samples')
  out.append((num(r[ie]), num(r[sm]), cur.split('/')[-1], ln, r[1][:110]))
tot, ts = sum(o[0] for o in out), sum(o[1] for o in out)
print('total warp instructions', tot, 'samples', ts)
for o in sorted(out, reverse=True)[:top]:
  print('%5.1f%% instr %5.1f%% samples  %s:%d  %s' % (100.*o[0]/tot, 100.*o[1]/max(ts, 1), o[2], o[3], o[4]))

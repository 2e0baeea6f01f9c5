#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/si
mkdir -p $O
echo "== pytest gpu (projection subset)"; timeout 600 python -m pytest tests -x -q -m gpu -k "project or links or pinned or golden or sparsevi or bpsvi" 2>&1 | tail -4
for f in 2 1 0; do
  echo "== K3 per-launch durations, BCG_PROJ_FAST=$f"
  BCG_PROJ_FAST=$f timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:project_ -c 5 --csv --log-file $O/k3_fast$f.csv python tools/omp_timing.py 1e6 512 1 > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('$O/k3_fast$f.csv')) if len(r)>10 and r[0].isdigit()]
print([ (r[4].split('(')[0][-30:], r[-1], r[-2]) for r in rows])
PY
done
echo "== projsum timing"; timeout 200 python tools/projsum_timing.py 1e6 200 512 2>&1 | tee $O/projsum.txt

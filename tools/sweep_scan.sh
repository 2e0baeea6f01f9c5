#!/bin/bash
# scan-geometry sweep on the GPU box: BCG_SCAN_WARPS x BCG_SCAN_STAGES x BCG_SCAN_STAGE_BYTES
W=${1:-lr_giga_N1e6_S256}
for cfg in "8 3 8192" "8 2 8192" "11 2 8192" "11 3 4096" "8 3 4096" "6 4 8192" "4 6 8192" "8 2 12288" "11 2 4096"; do
  set -- $cfg
  echo -n "warps=$1 stages=$2 stage_bytes=$3 : "
  BCG_SCAN_WARPS=$1 BCG_SCAN_STAGES=$2 BCG_SCAN_STAGE_BYTES=$3 timeout 300 python $(dirname $0)/../bench.py --workload $W --steps 100 --no-e2e --no-cpu-baseline --also none 2>&1 | tail -1 | python -c "
import sys,json
try:
  d=json.loads(sys.stdin.read()); print('%.1f it/s  %.4f ms/step  frac %.3f  %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel']))
except Exception as e: print('FAILED', e)
"
done

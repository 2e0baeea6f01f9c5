#!/bin/bash
# share of statically assigned chunks vs dynamic tail, at the per-GPU size of the 8-GPU run (N = 1.25e6, S = 512)
cd "$(dirname "$0")/.."
for pct in 100 95 90 80 50; do
  SWEEP_N=1.25e6 SWEEP_S=512 SWEEP_ALG=giga BCG_STATIC_PCT=$pct SWEEP_CFG="N=1.25e6 S=512 giga static_pct=$pct" timeout 200 python tools/filter_sweep.py one
done
SWEEP_N=1e7 SWEEP_S=512 SWEEP_ALG=giga BCG_STATIC_PCT=90 SWEEP_CFG="N=1e7 S=512 giga static_pct=90" timeout 200 python tools/filter_sweep.py one

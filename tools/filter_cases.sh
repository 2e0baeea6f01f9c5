#!/bin/bash
# float16 pre-filter on / off on the bench workloads (one process per case)
cd "$(dirname "$0")/.."
for case in "1e7 512 giga" "1e7 512 omp" "1e7 512 fw" "1e6 256 giga" "1e6 256 fw" "2e6 300 giga"; do
  set -- $case
  for f in 1 0; do
    SWEEP_N=$1 SWEEP_S=$2 SWEEP_ALG=$3 BCG_FILTER16=$f SWEEP_CFG="N=$1 S=$2 $3 filter16=$f" timeout 200 python tools/filter_sweep.py one
  done
done

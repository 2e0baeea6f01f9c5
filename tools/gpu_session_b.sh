#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/sb
mkdir -p $O
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/omp_launches.csv python tools/omp_timing.py 1e6 256 1 > $O/omp_launches.log 2>&1
tail -2 $O/omp_launches.log
wc -l $O/omp_launches.csv

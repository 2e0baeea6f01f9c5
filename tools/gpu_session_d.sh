#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/sd
mkdir -p $O
echo "== pytest gpu (omp/nnls subset)"; timeout 600 python -m pytest tests -x -q -m gpu -k "omp or nnls or golden or optimize" 2>&1 | tail -3
echo "== omp trace"; BCG_OMP_TRACE=1 timeout 300 python tools/omp_timing.py 1e6 256 1 2>&1 | tee $O/omp_c2_trace.txt
echo "== omp S512 trace"; BCG_OMP_TRACE=1 timeout 300 python tools/omp_timing.py 1e6 512 1 2>&1 | tee $O/omp_s512.txt

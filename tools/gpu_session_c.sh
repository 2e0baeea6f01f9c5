#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/sc
mkdir -p $O
echo "== pytest gpu"; (time timeout 900 python -m pytest tests -x -q -m gpu) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
echo "== omp timing downdate"; timeout 300 python tools/omp_timing.py 1e6 256 1 2>&1 | tee $O/omp_c2_downdate.txt
echo "== omp timing rebuild"; BCG_NNLS_DOWNDATE=0 timeout 300 python tools/omp_timing.py 1e6 256 1 2>&1 | tee $O/omp_c2_rebuild.txt
echo "== omp trace"; BCG_OMP_TRACE=1 timeout 300 python tools/omp_timing.py 1e6 256 1 2>&1 | tee $O/omp_c2_trace.txt
echo "== omp S512 N1e6"; timeout 300 python tools/omp_timing.py 1e6 512 1 2>&1 | tee $O/omp_s512.txt

#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/sf
mkdir -p $O
echo "== pytest gpu"; (time timeout 900 python -m pytest tests -x -q -m gpu) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
echo "== K3 timing"; timeout 300 python tools/project_timing.py 2e6 512 2>&1 | tee $O/k3_timing.jsonl
echo "== e2e breakdown"; timeout 300 python tools/e2e_breakdown.py > $O/e2e_breakdown.txt 2>&1; cat $O/e2e_breakdown.txt
echo "== ncu project fast"; timeout 300 ncu --set full --import-source on --clock-control none -k regex:project_fast_kernel -s 3 -c 1 -o $O/project_fast -f python tools/omp_timing.py 1e6 512 1 > $O/ncu_proj.log 2>&1; tail -2 $O/ncu_proj.log

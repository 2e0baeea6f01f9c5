#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/sh
mkdir -p $O
echo "== pytest gpu (projection subset)"; timeout 600 python -m pytest tests -x -q -m gpu -k "project or links or pinned or golden" 2>&1 | tail -4
echo "== K3 timing"; timeout 300 python tools/project_timing.py 2e6 512 2>&1 | tee $O/k3_timing.jsonl
echo "== ncu project pair"; timeout 300 ncu --set full --import-source on --clock-control none -k regex:project_pair_kernel -s 3 -c 1 -o $O/project_pair -f python tools/omp_timing.py 1e6 512 1 > $O/ncu_proj.log 2>&1; tail -2 $O/ncu_proj.log

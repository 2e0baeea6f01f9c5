#!/bin/bash
# one GPU-box session: parity tests, OMP / projection timings, bench, ncu captures (outputs under gpurun_out/)
cd "$(dirname "$0")/.."
O=gpurun_out/sa
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
echo "== pytest gpu"; (time timeout 900 python -m pytest tests -x -q -m gpu) > $O/pytest_gpu.log 2>&1; tail -5 $O/pytest_gpu.log
echo "== omp timing"; timeout 300 python tools/omp_timing.py 1e6 256 1,0 > $O/omp_c2.jsonl 2> $O/omp_c2.err; cat $O/omp_c2.jsonl; tail -3 $O/omp_c2.err
echo "== e2e breakdown"; timeout 300 python tools/e2e_breakdown.py > $O/e2e_breakdown.txt 2>&1; cat $O/e2e_breakdown.txt
echo "== projsum timing fast/slow"; timeout 200 python tools/projsum_timing.py 1e6 200 512 > $O/projsum_fast.txt 2>&1; BCG_FAST_LINK=0 timeout 200 python tools/projsum_timing.py 1e6 200 512 > $O/projsum_slow.txt 2>&1; cat $O/projsum_fast.txt $O/projsum_slow.txt
echo "== bench"; timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; cat $O/bench.json; tail -3 $O/bench.err
echo "== ncu omp"; timeout 300 ncu --set full --import-source on --clock-control none -k regex:omp_iteration_kernel -s 150 -c 1 -o $O/omp_iter -f python tools/omp_timing.py 1e6 256 1 > $O/ncu_omp.log 2>&1; tail -2 $O/ncu_omp.log
echo "== ncu project"; timeout 300 ncu --set full --import-source on --clock-control none -k regex:project_kernel -s 3 -c 1 -o $O/project_lr -f python tools/omp_timing.py 1e6 512 1 > $O/ncu_proj.log 2>&1; tail -2 $O/ncu_proj.log
ls -la $O

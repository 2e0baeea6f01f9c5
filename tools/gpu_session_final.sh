#!/bin/bash
# final session of the round: full GPU test suite, smoke, bench (1 GPU), secondary configs, ncu launch list + captures
cd "$(dirname "$0")/.."
O=gpurun_out/final
mkdir -p $O
echo "== pytest gpu"; (time timeout 900 python -m pytest tests -x -q -m gpu) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/smoke.txt
echo "== bench"; timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; cat $O/bench.json; tail -3 $O/bench.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 12 --warmup 3 > $O/bench_reference.json 2>> $O/bench.err; cat $O/bench_reference.json
echo "== configs c2 c3"; timeout 300 python bench_configs.py --config c2 2>&1 | grep "^{" | tee $O/configs.jsonl; timeout 300 python bench_configs.py --config c3 2>&1 | grep "^{" | tee -a $O/configs.jsonl
echo "== e2e breakdown"; timeout 300 python tools/e2e_breakdown.py > $O/e2e_breakdown.txt 2>&1; tail -6 $O/e2e_breakdown.txt
echo "== ncu launch list of the bench command"; timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/bench_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1; wc -l $O/bench_launches.csv
echo "== ncu project fast"; timeout 300 ncu --set full --import-source on --clock-control none -k regex:project_fast_kernel -s 3 -c 1 -o $O/project_fast -f python tools/omp_timing.py 1e6 512 1 > $O/ncu_proj.log 2>&1; tail -1 $O/ncu_proj.log
echo "== config c5"; timeout 400 python bench_configs.py --config c5 2>&1 | grep "^{" | tee -a $O/configs.jsonl

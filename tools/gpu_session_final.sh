#!/bin/bash
# round-end confirmation on one GPU: smoke, the whole GPU suite, the reference arm, the engine arm
cd "$(dirname "$0")/.."
O=gpurun_out/r02_final
mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 2>$O/ref.err | grep "^{" > $O/bench_reference.json; head -c 300 $O/bench_reference.json; echo
timeout 900 python bench.py --steps 20 --warmup 5 2>$O/bench.err | grep "^{" > $O/bench_1.json; head -c 600 $O/bench_1.json; echo
# launch list of the same bench command (per-launch times under ncu are cold-cache and serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/bench_launches.csv python bench.py --steps 20 --warmup 5 > $O/bench_under_ncu.log 2>&1; wc -l $O/bench_launches.csv

#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/sj
mkdir -p $O
echo "== pytest gpu"; (time timeout 900 python -m pytest tests -x -q -m gpu) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log | head -2
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout 600 python bench.py --no-cpu-baseline > $O/bench.json 2> $O/bench.err; python -c "
import json; d=json.load(open('$O/bench.json')); print(d['value'], d['e2e']['value'], d['e2e']['job'])"

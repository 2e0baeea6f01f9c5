#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/sg
mkdir -p $O
N=${1:-8}
echo "== c4 omp ($N)"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench_configs.py --config c4 2>&1 | grep "^{" | tee $O/c4_$N.jsonl

#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/se
mkdir -p $O
N=${1:-2}
echo "== mgpu check ($N)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/mgpu_check.py 2>&1 | grep -v Warning | tail -12 | tee $O/mgpu_check_$N.txt
echo "== c4 omp ($N)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench_configs.py --config c4 2>&1 | grep "^{" | tee $O/c4_$N.jsonl
echo "== bench ($N)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 200 --warmup 5 2>&1 | grep "^{" | tee $O/bench_$N.json

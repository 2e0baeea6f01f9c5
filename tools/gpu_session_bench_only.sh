#!/bin/bash
# bench.py exactly as the driver launches it on N GPUs (20 steps with the `also` block, then 200 steps without)
cd "$(dirname "$0")/.."
N=${1:-2}
O=gpurun_out/r02_bench_$N
mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 20 --warmup 5 2>$O/bench.err | grep "^{" > $O/bench_$N.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 200 --warmup 5 --also none 2>>$O/bench.err | grep "^{" > $O/bench_${N}_200.json
tail -c 200 $O/bench.err; ls -la $O

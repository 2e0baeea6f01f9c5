#!/bin/bash
# multi-GPU session: parity at every world size the box offers, the bench as the driver launches it, the loop timeline
cd "$(dirname "$0")/.."
N=${1:-2}
O=gpurun_out/r02_mgpu_$N
mkdir -p $O
echo "== pytest multi"; timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3 | tee $O/pytest_multi.txt
echo "== bench ($N)"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 20 --warmup 5 2>$O/bench.err | grep "^{" > $O/bench_$N.json; tail -c 300 $O/bench.err
echo "== bench 200 steps, no extras ($N)"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 200 --warmup 5 --also none 2>>$O/bench.err | grep "^{" > $O/bench_${N}_200.json
echo "== trace ($N)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/trace_loop.py lr_giga_N1e7_S512 60 2>/dev/null | grep -E "^\{|^ ?[0-9]|^it" > $O/trace_$N.txt; head -c 1500 $O/trace_$N.txt

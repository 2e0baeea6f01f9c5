#!/bin/bash
# float16 pre-filter: overflow test, compute-sanitizer on the small case (S = 256 runs the filtered kernels), ncu traffic capture
cd "$(dirname "$0")/.."
O=gpurun_out/f16
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "filter16_near" 2>&1 | tail -3
echo "== memcheck"; timeout 240 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanitizer_case.py 3000 > $O/memcheck.log 2>&1; grep -E "ERROR SUMMARY|SANITIZER CASE|Invalid|omp size" $O/memcheck.log | head
echo "== racecheck"; timeout 300 compute-sanitizer --tool racecheck --print-limit 5 python tools/sanitizer_case.py 1500 > $O/racecheck.log 2>&1; grep -E "RACECHECK SUMMARY|SANITIZER CASE|hazard|omp size" $O/racecheck.log | head
echo "== ncu traffic"; timeout 600 ncu --set full --clock-control none -k "regex:greedy_loop_kernel|omp_loop_kernel|scan_kernel" -o $O/r02_traffic_f16 -f python tools/ncu_traffic_case.py > $O/traffic.log 2>&1; grep -v "^==PROF" $O/traffic.log | tail -9; ls -la $O

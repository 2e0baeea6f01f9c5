#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/sk2
mkdir -p $O
echo "== memcheck"; timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanitizer_case.py 3000 > $O/memcheck.log 2>&1; grep -E "ERROR SUMMARY|SANITIZER CASE|Invalid|omp size" $O/memcheck.log | head
echo "== racecheck"; timeout 400 compute-sanitizer --tool racecheck --print-limit 5 python tools/sanitizer_case.py 1500 > $O/racecheck.log 2>&1; grep -E "RACECHECK SUMMARY|SANITIZER CASE|hazard|omp size" $O/racecheck.log | head

#!/bin/bash
# One GPU visit: tests, smoke, bench, launch list.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
bash tools/gpu_tests.sh 2>&1 | tee gpurun_out/tests.log | tail -25
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "=== bench"; timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench.log
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
wc -l gpurun_out/launches.csv

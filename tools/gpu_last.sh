#!/bin/bash
# Last pass of the round: the driver-style GPU suite (pair-GEMM tests included), smoke, the bench line, and one extra
# bench line with the CTA-pair GEMMs switched on (selection rule of launch_gemm).
mkdir -p gpurun_out
echo "=== GPU suite"; timeout 600 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/tests.log
echo "=== smoke"; timeout 200 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "=== bench"; timeout 400 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_n1.json | cut -c1-300
echo "=== bench with STTS_GEMM_2CTA=1"
STTS_GEMM_2CTA=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_pair.json | cut -c1-300

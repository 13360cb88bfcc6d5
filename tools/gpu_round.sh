#!/bin/bash
# Short validation pass on one B200: the driver-style GPU suite, smoke(), one bench line.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
echo "=== GPU suite"; timeout 600 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/tests.log
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4 | tee gpurun_out/smoke.log
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_n1.json | cut -c1-600

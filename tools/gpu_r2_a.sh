#!/bin/bash
# Round 2, first GPU pass: full GPU suite, smoke, short bench (both arms), compute-sanitizer memcheck / racecheck.
mkdir -p gpurun_out
echo "=== GPU suite"; timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/tests.log
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4 | tee gpurun_out/smoke.log
echo "=== bench"; timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_n1.json | cut -c1-600
echo "=== bench reference arm (3 steps)"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_reference_arm.json | cut -c1-400
echo "=== memcheck"; STTS_NO_GRAPH=1 timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/sanitizer_memcheck.log 2>&1; tail -5 gpurun_out/sanitizer_memcheck.log
echo "=== racecheck"; STTS_NO_GRAPH=1 timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/sanitizer_racecheck.log 2>&1; tail -5 gpurun_out/sanitizer_racecheck.log

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_chain.py -x -q 2>&1 | tail -3
timeout 300 python tools/trace_chain.py attn > gpurun_out/trace_attn.txt 2>&1
grep -E "item timeline|ms per|Error|kind 5|medians" gpurun_out/trace_attn.txt
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --in-flight 0 --no-config4 --no-other-precision"
for mode in 0 1; do
  echo "=== bench STTS_CHAIN_ATTN=$mode"
  STTS_CHAIN_ATTN=$mode timeout 600 $B 2>&1 | tail -1 > gpurun_out/bench_attn$mode.json
  python - <<PY
import json
j=json.loads(open("gpurun_out/bench_attn$mode.json").read().strip().splitlines()[-1]); print(j["ms_per_step"], j["stage_ms"], j.get("gpu_launches"))
PY
done

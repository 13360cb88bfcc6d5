#!/bin/bash
# in-chain attention: kernel tests, parity, A/B of the three launch schedules
mkdir -p gpurun_out
echo "=== chain tests"; timeout 900 python -m pytest tests/test_gpu_chain.py -x -q 2>&1 | tail -15
echo "=== parity subset (default schedule)"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "config2 or config3 or tight or denois or dit" 2>&1 | tail -5
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --in-flight 0 --no-config4 --no-other-precision"
for mode in 0 1 2; do
  echo "=== bench STTS_CHAIN_ATTN=$mode"
  STTS_CHAIN_ATTN=$mode timeout 600 $B 2>&1 | tail -1 > gpurun_out/bench_attn$mode.json
  python - <<PY
import json
j=json.loads(open("gpurun_out/bench_attn$mode.json").read().strip().splitlines()[-1]); print(j["ms_per_step"], j["stage_ms"], j.get("gpu_launches"))
PY
done

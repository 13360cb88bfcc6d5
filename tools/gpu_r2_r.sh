#!/bin/bash
mkdir -p gpurun_out
echo "=== vocoder tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -x -q -k "mix or convnext or vocoder or fused or head or config2_headline or ragged or edge or tight or encoder" 2>&1 | tail -3
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --in-flight 0 --no-config4 --no-other-precision"
for v in base; do
  timeout 600 $B 2>&1 | tail -1 > gpurun_out/bench_$v.json
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/bench_$v.json").read().strip().splitlines()[-1]); print("$v", j["ms_per_step"], j["stage_ms"], "tail", round(j["roofline"]["ms"],3), "front", round(j["roofline"]["front_ms"],3), j["roofline"]["frac"])
except Exception as e:
    print("$v failed", open("gpurun_out/bench_$v.json").read()[-500:])
PY
done

#!/bin/bash
mkdir -p gpurun_out
echo "=== fused tests"; timeout 90 python -m pytest tests/test_gpu_kernels.py -x -q -k "convnext_fused" 2>&1 | tail -2
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --in-flight 0 --no-config4 --no-other-precision"
for v in base gs1; do
  if [ $v = base ]; then L=""; else L=$PWD/smalltts_b200/variants/libsmalltts_b200_$v.so; fi
  STTS_LIB_PATH=$L timeout 120 $B 2>&1 | tail -1 > gpurun_out/bench_$v.json
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/bench_$v.json").read().strip().splitlines()[-1]); print("$v", j["ms_per_step"], "tail", round(j["roofline"]["ms"],3), "front", round(j["roofline"]["front_ms"],3), round(j["roofline"]["frac"],4))
except Exception as e:
    print("$v failed", open("gpurun_out/bench_$v.json").read()[-300:])
PY
done

#!/bin/bash
mkdir -p gpurun_out
echo "=== vocoder tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -x -q -k "convnext or vocoder or fused or config2_headline or ragged" 2>&1 | tail -3
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --in-flight 0 --no-config4 --no-other-precision"
timeout 600 $B 2>&1 | tail -1 > gpurun_out/bench_base.json
python - <<PY
import json
j=json.loads(open("gpurun_out/bench_base.json").read().strip().splitlines()[-1]); print(j["ms_per_step"], j["stage_ms"], "tail", round(j["roofline"]["ms"],3), "front", round(j["roofline"]["front_ms"],3), j["roofline"]["frac"])
PY
L=$PWD/smalltts_b200/variants/libsmalltts_b200_ftrace.so
for c in 32 64; do echo "=== C=$c"; STTS_LIB_PATH=$L timeout 300 python tools/trace_fused.py $c 2>&1 | tail -15 | head -6; done

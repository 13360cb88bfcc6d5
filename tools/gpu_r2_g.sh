#!/bin/bash
mkdir -p gpurun_out
echo "=== kernel tests (fused)"; timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "fused or convnext" 2>&1 | tail -3
L=$PWD/smalltts_b200/variants/libsmalltts_b200_ftrace.so
echo "=== fused trace C=64"; STTS_LIB_PATH=$L timeout 300 python tools/trace_fused.py 64 2>&1 | tail -16 | tee gpurun_out/trace_fused64.txt
echo "=== fused trace C=32"; STTS_LIB_PATH=$L timeout 300 python tools/trace_fused.py 32 2>&1 | tail -5
echo "=== parity subset"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "vocoder or config2_headline or encoder or tight" 2>&1 | tail -3
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --in-flight 0 --no-config4 --no-other-precision"
echo "=== bench"; timeout 600 $B 2>&1 | tail -1 > gpurun_out/bench_out8.json
python - <<'PY'
import json
j=json.loads(open("gpurun_out/bench_out8.json").read().strip().splitlines()[-1]); print(j["ms_per_step"], j["stage_ms"], j["roofline"]["ms"], j["roofline"]["frac"], j["roofline"]["front_ms"])
PY

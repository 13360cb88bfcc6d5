#!/bin/bash
mkdir -p gpurun_out
echo "=== tight parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -s -k "tight or config2_headline" 2>&1 | grep -E "tight|passed|failed|Error|assert" | tail -30 | tee gpurun_out/tight.log
echo "=== bench (both precisions)"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --in-flight 0 --no-config4 2>&1 | tail -1 > gpurun_out/bench_prec.json
python - <<'PY'
import json
j=json.loads(open("gpurun_out/bench_prec.json").read().strip().splitlines()[-1]); print(j["ms_per_step"], j["stage_ms"], j.get("other_precision"))
PY

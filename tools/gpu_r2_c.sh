#!/bin/bash
mkdir -p gpurun_out
echo "=== chain kernel tests"; timeout 600 python -m pytest tests/test_gpu_chain.py -x -q 2>&1 | tail -5
echo "=== trace fused"; timeout 300 python tools/trace_chain.py 2>&1 | tail -60 | tee gpurun_out/trace_fused.txt
echo "=== parity (chain path)"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "config or fixture or ragged or edge or teacher" 2>&1 | tail -5
echo "=== bench chain"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --in-flight 0 --no-config4 2>&1 | tail -1 > gpurun_out/bench_chain.json
echo "=== bench chain split"; STTS_CHAIN_SPLIT=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --in-flight 0 --no-config4 2>&1 | tail -1 > gpurun_out/bench_chain_split.json
python - <<'PY'
import json
for n in ("bench_chain","bench_chain_split"):
    try:
        j=json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1]); print(n, j["ms_per_step"], j["stage_ms"])
    except Exception as e: print(n, "failed", e)
PY

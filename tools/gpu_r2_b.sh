#!/bin/bash
# Round 2, chained DiT kernel bring-up: kernel tests, path parity, A/B/C bench (chain | chain split per GEMM | generic).
mkdir -p gpurun_out
echo "=== chain kernel tests"; timeout 600 python -m pytest tests/test_gpu_chain.py -x -q -s 2>&1 | tail -40 | tee gpurun_out/chain_tests.log
echo "=== parity (chain path)"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -15 | tee gpurun_out/parity_chain.log
echo "=== bench chain"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --in-flight 0 --no-config4 2>&1 | tail -1 | tee gpurun_out/bench_chain.json | cut -c1-300
echo "=== bench chain split"; STTS_CHAIN_SPLIT=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --in-flight 0 --no-config4 2>&1 | tail -1 | tee gpurun_out/bench_chain_split.json | cut -c1-300
echo "=== bench generic"; STTS_NO_CHAIN=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --in-flight 0 --no-config4 2>&1 | tail -1 | tee gpurun_out/bench_nochain.json | cut -c1-300
python - <<'PY'
import json
for n in ("bench_chain","bench_chain_split","bench_nochain"):
    try:
        j=json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1]); print(n, j["ms_per_step"], j["stage_ms"])
    except Exception as e: print(n, "failed", e)
PY

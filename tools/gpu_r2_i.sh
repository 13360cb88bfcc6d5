#!/bin/bash
mkdir -p gpurun_out
echo "=== split-K kernel tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "split_k or linear" 2>&1 | tail -4
echo "=== parity subset"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "vocoder or config2_headline or encoder or tight or clone" 2>&1 | tail -3
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --in-flight 0 --no-config4 --no-other-precision"
echo "=== bench split-K"; timeout 600 $B 2>&1 | tail -1 > gpurun_out/bench_split.json
echo "=== bench no split"; STTS_NO_SPLITK=1 timeout 600 $B 2>&1 | tail -1 > gpurun_out/bench_nosplit.json
python - <<'PY'
import json
for n in ("bench_split","bench_nosplit"):
    j=json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1]); print(n, j["ms_per_step"], j["stage_ms"], "tail", j["roofline"]["ms"], "front", j["roofline"]["front_ms"])
PY

#!/bin/bash
mkdir -p gpurun_out
echo "=== chain kernel tests"; timeout 600 python -m pytest tests/test_gpu_chain.py -x -q 2>&1 | tail -3
echo "=== trace fused (kpb2, counter polling)"; timeout 300 python tools/trace_chain.py 2>&1 | grep -E "phase|medians|MMAs issued|counted|flag seen" | tee gpurun_out/trace_fused.txt
echo "=== trace fused kpb1"; STTS_LIB_PATH=$PWD/smalltts_b200/variants/libsmalltts_b200_kpb1.so timeout 300 python tools/trace_chain.py 2>&1 | grep -E "phase|medians|MMAs issued|counted|flag seen" | tee gpurun_out/trace_fused_kpb1.txt
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --in-flight 0 --no-config4"
echo "=== bench chain"; timeout 600 $B 2>&1 | tail -1 > gpurun_out/bench_chain.json
echo "=== bench chain kpb1"; STTS_LIB_PATH=$PWD/smalltts_b200/variants/libsmalltts_b200_kpb1.so timeout 600 $B 2>&1 | tail -1 > gpurun_out/bench_chain_kpb1.json
echo "=== bench chain, attention skipped"; STTS_DEBUG_SKIP_ATTENTION=1 timeout 600 $B 2>&1 | tail -1 > gpurun_out/bench_chain_noattn.json
python - <<'PY'
import json
for n in ("bench_chain","bench_chain_kpb1","bench_chain_noattn"):
    try:
        j=json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1]); print(n, j["ms_per_step"], j["stage_ms"])
    except Exception as e: print(n, "failed", e)
PY

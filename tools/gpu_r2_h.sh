#!/bin/bash
mkdir -p gpurun_out
echo "=== tight + cpp host tests"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cpp_host.py -x -q -s -k "tight or vocoder_vs or cpp" 2>&1 | grep -E "tight|C\+\+|passed|failed|Error|assert" | tail -20
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --in-flight 0 --no-config4 --no-other-precision"
for v in "" hint2000 hint20000; do
  if [ -n "$v" ]; then export STTS_LIB_PATH=$PWD/smalltts_b200/variants/libsmalltts_b200_$v.so; fi
  echo "=== bench variant '$v'"; timeout 600 $B 2>&1 | tail -1 > gpurun_out/bench_v_$v.json
  python - "$v" <<'PY'
import json,sys
j=json.loads(open(f"gpurun_out/bench_v_{sys.argv[1]}.json").read().strip().splitlines()[-1]); print(sys.argv[1], j["ms_per_step"], j["stage_ms"], "tail", j["roofline"]["ms"], "front", j["roofline"]["front_ms"])
PY
done

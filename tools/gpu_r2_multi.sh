#!/bin/bash
# Multi-GPU pass: N = $1 ranks.  The 2-GPU single-process test, weak-scaling headline, configs[3] strong scaling.
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  echo "=== devices=[0,1] test"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "devices_kwarg or engine_clone" 2>&1 | tail -3
fi
echo "=== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/bench_n$N.err | tail -1 > gpurun_out/bench_n$N.json
python - $N <<'PY'
import json,sys
n=sys.argv[1]
try:
    j=json.loads(open(f"gpurun_out/bench_n{n}.json").read().strip().splitlines()[-1]); print("N",n,"value",j["value"],"ms/step",j["ms_per_step"],"e2e",j["e2e"]["value"]); print("config4",j.get("config4"))
except Exception as e:
    print("failed",e); print(open(f"gpurun_out/bench_n{n}.err").read()[-2000:])
PY

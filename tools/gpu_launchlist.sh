#!/bin/bash
mkdir -p gpurun_out
STTS_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^(?!.*pack_).*" -c 4000 --csv --log-file gpurun_out/launches.csv \
  python tools/profile_synth.py 2 > gpurun_out/ncu_synth.log 2>&1
tail -1 gpurun_out/ncu_synth.log; wc -l gpurun_out/launches.csv

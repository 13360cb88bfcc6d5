#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"${1:-convnext_fused}" \
  --launch-skip ${2:-6} --launch-count 1 -o /tmp/fused -f python tools/profile_decode.py 2 > gpurun_out/prof_fused.log 2>&1
tail -1 gpurun_out/prof_fused.log
cp /tmp/fused.ncu-rep gpurun_out/fused.ncu-rep
ncu -i /tmp/fused.ncu-rep --page raw --csv > gpurun_out/fused_raw.csv 2>/dev/null
ls -la /tmp/fused.ncu-rep gpurun_out/fused_source.csv gpurun_out/fused_raw.csv

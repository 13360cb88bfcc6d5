#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"convnext_fused" --launch-skip 2 --launch-count 3 \
  -o gpurun_out/fused -f python tools/profile_decode.py 2 > gpurun_out/prof_fused.log 2>&1
tail -2 gpurun_out/prof_fused.log; ls -la gpurun_out/fused.ncu-rep

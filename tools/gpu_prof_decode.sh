#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout 900 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section Occupancy --section LaunchStats \
  --clock-control none -k regex:"gemm_kernel|convnext_mix|head_conv" \
  --launch-skip 86 --launch-count 86 -o /tmp/decode_sections -f python tools/profile_decode.py 2 > gpurun_out/prof_decode.log 2>&1
tail -2 gpurun_out/prof_decode.log
ncu -i /tmp/decode_sections.ncu-rep --page raw --csv > gpurun_out/decode_sections_raw.csv 2>/dev/null
ls -la /tmp/decode_sections.ncu-rep gpurun_out/decode_sections_raw.csv

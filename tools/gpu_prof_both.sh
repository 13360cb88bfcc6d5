#!/bin/bash
# ncu --set full captures: one DiT block (7 kernels) and one fused ConvNeXt layer; reps land in gpurun_out/
mkdir -p gpurun_out
STTS_NO_GRAPH=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:"attention_kernel|gemm_kernel|row_norm_kernel" \
  --launch-skip ${1:-705} --launch-count ${2:-7} -o gpurun_out/dit -f python tools/profile_synth.py 2 > gpurun_out/prof_dit.log 2>&1
tail -1 gpurun_out/prof_dit.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"convnext_fused" \
  --launch-skip 6 --launch-count 1 -o gpurun_out/fused -f python tools/profile_decode.py 2 > gpurun_out/prof_fused.log 2>&1
tail -1 gpurun_out/prof_fused.log
ls -la gpurun_out/*.ncu-rep

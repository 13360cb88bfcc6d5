#!/bin/bash
mkdir -p gpurun_out
# one DiT block of the second eager synthesize: row_norm, gemm(qkvg), attention, gemm(O), row_norm, gemm(swiglu), gemm(W2)
STTS_NO_GRAPH=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"attention_kernel|gemm_kernel|row_norm_kernel" \
  --launch-skip ${1:-705} --launch-count ${2:-7} -o /tmp/dit -f python tools/profile_synth.py 2 > gpurun_out/prof_dit.log 2>&1
tail -1 gpurun_out/prof_dit.log
cp /tmp/dit.ncu-rep gpurun_out/dit.ncu-rep; ls -la gpurun_out/dit.ncu-rep

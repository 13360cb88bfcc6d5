#!/bin/bash
# Run the GPU test-suite in isolated processes (a CUDA fault poisons the context of its process only).
mkdir -p gpurun_out
for k in "linear or small_k or swiglu or epilogue or causal or transpose" grouped attention convnext_mix convnext_fused ffn_fused; do
  echo "=== kernels: $k"
  timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "$k" 2>&1 | grep -E "passed|failed|Error|error|assert|FAILED" | head -30
done
echo "=== parity"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | grep -v "^$" | head -150 > gpurun_out/parity.log
grep -E "passed|failed|Error|assert|FAILED|rel_l2|E  " gpurun_out/parity.log | head -60

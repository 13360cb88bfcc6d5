#!/bin/bash
# CTA-pair GEMM experiment: kernel tests, then path parity and one bench line with the pair variant switched on.
mkdir -p gpurun_out
echo "=== pair kernel tests"; STTS_TEST_PAIR=1 timeout 150 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "cta_pair" 2>&1 | tail -12 | tee gpurun_out/pair_tests.log
if grep -q " passed" gpurun_out/pair_tests.log && ! grep -q "failed\|error" gpurun_out/pair_tests.log; then
  echo "=== path parity with STTS_GEMM_2CTA=1"
  STTS_GEMM_2CTA=1 timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "vocoder_vs_reference or config1_end_to_end or full_size or codec_encoder_vs_reference" 2>&1 | tail -5 | tee gpurun_out/pair_parity.log
  echo "=== bench with STTS_GEMM_2CTA=1"
  STTS_GEMM_2CTA=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_pair.json | cut -c1-900
fi

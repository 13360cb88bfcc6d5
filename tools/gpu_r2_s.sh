#!/bin/bash
echo "== setmaxnreg: launch 80, GELU 64, out 112"; STTS_LIB_PATH=$PWD/smalltts_b200/variants/libsmalltts_b200_smr64.so timeout 60 python -m pytest tests/test_gpu_kernels.py -x -q -k "convnext_fused" 2>&1 | tail -2

#!/bin/bash
echo "== default (MG=2, NXB=3 for C=32)"; timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "convnext_fused" 2>&1 | tail -3
for i in 1 2 3; do timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "convnext_fused" 2>&1 | tail -1; done

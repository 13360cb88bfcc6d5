#!/bin/bash
mkdir -p gpurun_out
for v in ftrace; do
L=$PWD/smalltts_b200/variants/libsmalltts_b200_$v.so
echo "=== $v C=32"; STTS_LIB_PATH=$L timeout 300 python tools/trace_fused.py 32 2>&1 | tail -26 | tee gpurun_out/trace_${v}_32.txt
done

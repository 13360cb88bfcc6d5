#!/bin/bash
L=$PWD/smalltts_b200/variants/libsmalltts_b200_ftrace.so
for c in 32 64; do echo "=== C=$c"; STTS_LIB_PATH=$L timeout 120 python tools/trace_fused.py $c 2>&1 | tail -22 | head -9; done

#!/bin/bash
# memcheck over every kernel family with the round's new code paths: default schedule and the one-launch chained kernel
mkdir -p gpurun_out
echo "=== memcheck (default schedule)"; STTS_NO_GRAPH=1 timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/sanitizer_memcheck.log 2>&1; tail -4 gpurun_out/sanitizer_memcheck.log
echo "=== memcheck (STTS_CHAIN_ATTN=1)"; STTS_CHAIN_ATTN=1 STTS_NO_GRAPH=1 timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/sanitizer_memcheck_chain_attn.log 2>&1; tail -4 gpurun_out/sanitizer_memcheck_chain_attn.log

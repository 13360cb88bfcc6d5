#!/bin/bash
# Round-end measurement pass on one B200: tests, smoke, bench (both arms), launch list, per-kernel DRAM bytes of a
# decode, ncu --set full of the top kernels.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
echo "=== driver-style single-process GPU suite"; timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/tests.log
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4 | tee gpurun_out/smoke.log
echo "=== bench"; timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_n1.json | cut -c1-400
echo "=== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_reference_arm.json | cut -c1-300
echo "=== launch list"
STTS_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^(?!.*pack_).*" -c 4000 --csv --log-file gpurun_out/launches.csv \
  python tools/profile_synth.py 2 > gpurun_out/ncu_synth.log 2>&1
tail -1 gpurun_out/ncu_synth.log | cut -c1-200; wc -l gpurun_out/launches.csv
echo "=== decode dram bytes"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"gemm_kernel|convnext|head_conv|ffn_fused" \
  --csv --log-file gpurun_out/decode_dram.csv python tools/profile_decode.py 2 > gpurun_out/decode_dram.log 2>&1
tail -1 gpurun_out/decode_dram.log; wc -l gpurun_out/decode_dram.csv
echo "=== ncu full: fused tail layer, attention, DiT GEMMs"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"convnext_fused|ffn_fused" --launch-skip 8 --launch-count 2 \
  -o gpurun_out/fused -f python tools/profile_decode.py 2 > gpurun_out/prof_fused.log 2>&1
STTS_NO_GRAPH=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:"attention_kernel|gemm_kernel|row_norm_kernel|head_split" \
  --launch-skip ${1:-787} --launch-count 8 -o gpurun_out/dit -f python tools/profile_synth.py 2 > gpurun_out/prof_dit.log 2>&1
ls -la gpurun_out/*.ncu-rep

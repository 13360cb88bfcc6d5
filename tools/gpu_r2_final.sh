#!/bin/bash
# Round-2 measurement pass on one B200: full GPU suite, smoke, bench (both arms), launch list of one synthesize,
# per-kernel DRAM bytes of a decode, ncu --set full of the chained DiT kernel / attention / fused tail kernels / front GEMMs.
mkdir -p gpurun_out
echo "=== GPU suite"; timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/tests.log
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4 | tee gpurun_out/smoke.log
echo "=== bench"; timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_n1.json | cut -c1-300
echo "=== bench reference arm"; timeout 900 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_reference_arm.json | cut -c1-300
echo "=== launch list (second eager synthesize)"
STTS_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^(?!.*pack_).*" -c 6000 --csv --log-file gpurun_out/launches.csv \
  python tools/profile_synth.py 2 > gpurun_out/ncu_synth.log 2>&1
tail -1 gpurun_out/ncu_synth.log | cut -c1-200; wc -l gpurun_out/launches.csv
echo "=== decode dram bytes"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"gemm_kernel|convnext|head_conv|ffn_fused" \
  --csv --log-file gpurun_out/decode_dram.csv python tools/profile_decode.py 2 > gpurun_out/decode_dram.log 2>&1
tail -1 gpurun_out/decode_dram.log; wc -l gpurun_out/decode_dram.csv
echo "=== ncu full: chained DiT kernel + attention (second synthesize, block 5)"
STTS_NO_GRAPH=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:"dit_chain_kernel|attention_kernel" \
  --launch-skip 70 --launch-count 4 -o gpurun_out/dit_chain -f python tools/profile_synth.py 2 > gpurun_out/prof_dit.log 2>&1
echo "=== ncu full: fused tail kernels + head conv"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"convnext_fused|ffn_fused|head_conv|convnext_mix_rows" --launch-skip 16 --launch-count 7 \
  -o gpurun_out/tail -f python tools/profile_decode.py 2 > gpurun_out/prof_tail.log 2>&1
ls -la gpurun_out/*.ncu-rep

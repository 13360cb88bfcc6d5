#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum --clock-control none -k regex:"gemm_kernel|convnext|head_conv" \
  --launch-skip ${1:-74} --launch-count ${2:-80} --csv --log-file gpurun_out/decode_times.csv python tools/profile_decode.py 2 > gpurun_out/decode_times.log 2>&1
tail -1 gpurun_out/decode_times.log

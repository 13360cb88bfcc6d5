#!/bin/bash
for v in base mg1 nosmr; do
  if [ $v = base ]; then L=""; else L=$PWD/smalltts_b200/variants/libsmalltts_b200_$v.so; fi
  for i in 1 2 3; do echo -n "$v run $i: "; STTS_LIB_PATH=$L timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "full_size_config2" 2>&1 | tail -1; done
done

#!/bin/bash
for v in base nxb2 nosmr; do
  if [ $v = base ]; then L=""; else L=$PWD/smalltts_b200/variants/libsmalltts_b200_$v.so; fi
  echo -n "$v: "; STTS_LIB_PATH=$L timeout 100 python tools/bench_concurrent.py 10 2 0 2>/dev/null | grep "^{" | python -c "
import sys,json
print([ (json.loads(l)['batches_in_flight'], round(json.loads(l)['value'])) for l in sys.stdin])"
done

#!/usr/bin/env python
"""Write the seeded random weights of smalltts_b200.synthetic as `.sttsw` containers (dit / decoder / encoder) for the
compiled callers (examples/bench_pipeline.cpp): no checkpoint can be fetched offline.

    python tools/make_synthetic_sttsw.py OUTDIR [--bf16]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smalltts_b200 import synthetic, weights

out = sys.argv[1] if len(sys.argv) > 1 else "."
dtype = "bfloat16" if "--bf16" in sys.argv else "float32"
os.makedirs(out, exist_ok=True)
for name, sd in (("dit", synthetic.dit_state_dict(0)), ("decoder", synthetic.vocoder_state_dict(1)),
                 ("encoder", synthetic.encoder_state_dict(2))):
    path = os.path.join(out, name + ".sttsw")
    weights.save_packed(path, sd, dtype=dtype)
    print(path, f"{os.path.getsize(path) / 1e6:.0f} MB")

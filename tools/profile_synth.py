#!/usr/bin/env python
"""One eager synthesize of BASELINE configs[1] (B=8 x 10 s) for ncu launch lists.  Run with STTS_NO_GRAPH=1."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smalltts_b200 import synthetic
from smalltts_b200.engine import Engine, pad_batch

eng = Engine(0)
eng.load_state_dicts(synthetic.dit_state_dict(0), synthetic.vocoder_state_dict(1))
refs, ids, frames, _ = synthetic.synthetic_inputs(8, 75, 15, 120)
ref, ref_len, idt, ph_len = pad_batch(refs, ids, frames)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    a = eng.synthesize(ref, ref_len, idt, ph_len, frames, 75, seed=1 + i)
print("ok", a.shape, eng.timings())

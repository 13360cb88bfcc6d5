#!/usr/bin/env python
"""One vocoder decode (B=8, T=75) through the C ABI, for ncu.  usage: profile_decode.py [passes]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from smalltts_b200 import synthetic
from smalltts_b200.engine import Engine

eng = Engine(0)
eng.load_state_dicts(synthetic.dit_state_dict(0), synthetic.vocoder_state_dict(1))
lat = np.random.default_rng(0).standard_normal((8, 75, 64)).astype(np.float32)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    a = eng.decode(lat)
print("decode ok", a.shape, float(np.abs(a).mean()), eng.vocoder_ms())

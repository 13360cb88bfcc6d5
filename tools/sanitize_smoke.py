#!/usr/bin/env python
"""Small eager pass through every kernel family of the engine (condition encoder, DMD loop, vocoder with the fused
C <= 128 kernels, codec encoder, resampler) for compute-sanitizer:

    STTS_NO_GRAPH=1 compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
    STTS_NO_GRAPH=1 compute-sanitizer --tool racecheck python tools/sanitize_smoke.py

Shapes are tiny but ragged (partial tiles, masked rows), so out-of-bounds and hazard reports are about the real
code paths; the logs are kept under profiles/."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from smalltts_b200 import synthetic
from smalltts_b200.infer import SmallTTS

tts = SmallTTS(state_dicts=(synthetic.dit_state_dict(0), synthetic.vocoder_state_dict(1), synthetic.encoder_state_dict(2)))
refs, ids, frames, noise = synthetic.synthetic_inputs(3, [5, 2, 7], [4, 6, 3], [9, 5, 12], seed=3)
durs = [f * 3200 / 24000 + 1e-3 for f in frames]
out = tts.synthesize_batch(refs, ids, durs, noise=noise.numpy())
assert all(np.isfinite(a).all() for a in out)
out = tts.synthesize_batch(refs, ids, durs, seed=5)  # Philox path
lat = tts.engine.encode_audio(np.zeros((1, 2 * 3200), np.float32) + 0.1)
y = tts.engine.resample(np.random.default_rng(0).standard_normal((1, 4410)).astype(np.float32), 44100, 24000)
print("sanitize_smoke ok", [a.shape for a in out], lat.shape, y.shape)

#!/usr/bin/env python
"""BASELINE configs[4]: teacher 128-step flow ODE (3-way CFG, DDIM) vs DMD 4-step on the same prompt set, one B200.
Latency per batch (CUDA events on the engine stream) + how far the student's latents are from the teacher's
(relative L2; random-init weights, so this is a plumbing number, not a quality claim).
usage: sweep_teacher.py [steps ...]  -> one JSON line per sampler"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from smalltts_b200 import synthetic
from smalltts_b200.engine import Engine, pad_batch

B, T, R, P = 8, 75, 15, 120
eng = Engine(0)
eng.load_state_dicts(synthetic.dit_state_dict(0), synthetic.vocoder_state_dict(1))
refs, ids, frames, noise = synthetic.synthetic_inputs(B, T, R, P)
ref, ref_len, idt, ph_len = pad_batch(refs, ids, frames)
audio_s = B * T * 3200 / 24000
x1 = noise[0].numpy()


def timed(fn, reps=3):
    fn()
    eng.timer_start()
    for _ in range(reps):
        out = fn()
    return eng.timer_stop() / reps, out


cond = eng.encode_conditions(ref, ref_len, idt, ph_len)
ms, lat_dmd = timed(lambda: eng.sample(cond, frames, T, noise=noise.numpy(), steps=4))
ms_dec, _ = timed(lambda: eng.decode(lat_dmd))
print(json.dumps({"sampler": "dmd", "steps": 4, "evals_per_utt": 4, "sample_ms": ms, "decode_ms": ms_dec,
                  "rtf": (ms + ms_dec) / 1e3 / audio_s}))
cond3 = eng.encode_conditions_cfg(ref, ref_len, idt, ph_len)
for steps in [int(a) for a in sys.argv[1:]] or [16, 128]:
    ms, lat = timed(lambda: eng.sample_teacher(cond3, frames, T, steps=steps, noise=x1), reps=1 if steps > 32 else 3)
    rel = float(np.linalg.norm(lat - lat_dmd) / np.linalg.norm(lat))
    print(json.dumps({"sampler": "teacher_ddim_cfg3", "steps": steps, "evals_per_utt": 3 * steps, "sample_ms": ms,
                      "decode_ms": ms_dec, "rtf": (ms + ms_dec) / 1e3 / audio_s, "finite": bool(np.isfinite(lat).all()),
                      "student_vs_teacher_rel_l2": rel}))

#!/usr/bin/env python
"""BASELINE configs[2]: voice cloning -- ONE reference wav + 16 target prompts on one B200 (scripts/infer/clone.py path).

    python tools/bench_config3.py [src_sample_rate]           # default 44100: the resampler is part of the path

SURVEY 8(d) C3 inputs: a 3 s 440 Hz sine reference (src/server/src/bin/bench.rs:8-13), 16 prompts with T ~ U{15..75}
frames and P = round(1.53 T) phonemes.  Timed: (a) `clone_voice` = H2D + mono mix + resample_hq + codec encoder + D2H,
once per voice; (b) the 16 prompts as length-bucketed micro-batches that all share the cloned voice.  Prints one JSON
line (wall clock, best of 3 after a warm-up pass that builds the per-shape plans)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from smalltts_b200 import parallel, synthetic
from smalltts_b200.infer import SmallTTS

sr = int(sys.argv[1]) if len(sys.argv) > 1 else 44100
rng = np.random.default_rng(20260217)
N = 16
frames = rng.integers(15, 76, N).tolist()
phon_n = [int(round(1.53 * f)) for f in frames]
_, ids, _, _ = synthetic.synthetic_inputs(N, frames, 1, phon_n, steps=1)
durs = [f * 3200 / 24000 + 1e-3 for f in frames]
wav = (0.5 * np.sin(2 * np.pi * 440.0 * np.arange(3 * sr) / sr)).astype(np.float32)
tts = SmallTTS.synthetic(encoder_seed=2)


def clone():
    return tts.clone_voice(wav, sample_rate=sr)


def prompts(ref):
    order = sorted(range(N), key=lambda i: -frames[i])
    out = [None] * N
    for mb in parallel.length_buckets(order, frames):
        for i, a in zip(mb, tts.synthesize_batch([ref] * len(mb), [ids[i] for i in mb], [durs[i] for i in mb], seed=7)):
            out[i] = a
    return out


best_clone = best_syn = None
for rep in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ref = clone()
    t1 = time.perf_counter()
    out = prompts(ref)
    t2 = time.perf_counter()
    if rep > 0:
        best_clone = (t1 - t0) if best_clone is None else min(best_clone, t1 - t0)
        best_syn = (t2 - t1) if best_syn is None else min(best_syn, t2 - t1)
audio_s = sum(frames) * 3200 / 24000
assert ref.shape == (22, 64) and all(a.shape == (1, f * 3200) and np.isfinite(a).all() for a, f in zip(out, frames))
print(json.dumps({"workload": f"configs[2]: 1 reference wav (3 s, {sr} Hz) + 16 prompts, one B200",
                  "clone_voice_ms": best_clone * 1e3, "codec_enc_ms": tts.engine.timings()["codec_enc_ms"],
                  "prompts_ms": best_syn * 1e3, "audio_seconds": audio_s,
                  "audio_s_per_s": audio_s / (best_clone + best_syn), "rtf": (best_clone + best_syn) / audio_s,
                  "micro_batches": len(parallel.length_buckets(sorted(range(N), key=lambda i: -frames[i]), frames)),
                  "timing": "wall clock incl. H2D/D2H; best of 3 after a warm-up pass"}))

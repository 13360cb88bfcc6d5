"""The reference's scripts/tryme.py on the B200 engine: default voice latents + one sentence -> out/tryme.wav.

    python scripts/tryme.py "some text" [--latents assets/tryme/latents.npy] [--tokens 1,2,3] [model args]
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from clone import add_model_args, load_tts, tokens_for  # noqa: E402

from smalltts_b200.infer import estimate_duration  # noqa: E402
from smalltts_b200.serve import encode_wav  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("text", nargs="?",
                    default="hello this is small brain speaking, thanks for trying this model out and have fun")
    ap.add_argument("--latents", default="assets/tryme/latents.npy")
    ap.add_argument("--tokens", default="")
    add_model_args(ap)
    args = ap.parse_args()
    os.makedirs("out", exist_ok=True)
    print("loading model")
    model = load_tts(args)
    if os.path.exists(args.latents):
        ref_latents = np.load(args.latents).astype(np.float32)
    elif args.synthetic:
        ref_latents = np.random.default_rng(0).standard_normal((15, 64)).astype(np.float32)
    else:
        raise FileNotFoundError(args.latents)
    tokens = tokens_for(args.text, args.tokens)
    duration = estimate_duration(args.text)
    print(f"generating ({duration:.1f}s estimated)")
    audio = model.synthesize(ref_latents, tokens, duration)
    with open("out/tryme.wav", "wb") as fh:
        fh.write(encode_wav(audio.squeeze(), 24_000))
    print("out/tryme.wav")

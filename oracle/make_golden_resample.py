"""Generate tests/golden/resample_small.npz by running the reference's own ``resample_hq`` (test infrastructure).

    python oracle/make_golden_resample.py

``/root/reference/src/smalltts/infer/utils.py:7-23`` (torchaudio Resample, sinc_interp_kaiser, lowpass_filter_width
1024, rolloff 0.94) is imported as it is and run on short seeded signals at the sample rates a reference wav is likely
to come in (scripts/infer/clone.py:27-32 resamples whatever ``torchaudio.load`` returns to 24 kHz).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, "/root/reference/src")

OUT = os.path.join(ROOT, "tests", "golden")


def main() -> None:
    from smalltts.infer.utils import resample_hq  # the reference itself

    g = torch.Generator().manual_seed(4242)
    out = {}
    for sr, n in ((44100, 9000), (48000, 7001), (16000, 5000), (22050, 6000), (8000, 2500), (32000, 4097)):
        t = torch.arange(n) / sr
        x = 0.4 * torch.sin(2 * torch.pi * 440.0 * t)[None] + 0.2 * torch.randn(2, n, generator=g)  # (2, n)
        y = resample_hq(x, sr, 24000)
        out[f"x_{sr}"] = x.numpy()
        out[f"y_{sr}"] = y.numpy()
        print(sr, tuple(x.shape), "->", tuple(y.shape), float(y.abs().max()))
    np.savez_compressed(os.path.join(OUT, "resample_small.npz"), **out)


if __name__ == "__main__":
    main()

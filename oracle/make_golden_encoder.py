"""Generate tests/golden/encoder_small.npz with transformers' VibeVoiceAcousticTokenizerEncoderModel (test infrastructure).

    python oracle/make_golden_encoder.py

SURVEY 8(a19): the reference's codec encoder is the un-vendored ``assets/codec/encoder.onnx`` (codec/onnx.py:56-75), an
export of microsoft/VibeVoice's acoustic tokenizer; its published arithmetic is transformers 5.5.0
``VibeVoiceAcousticTokenizerEncoderModel`` (hf:300-403).  Seeded weights from smalltts_b200.synthetic.encoder_state_dict.
Against the ONNX export itself: parity unpinned (no file, no network).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from smalltts_b200 import synthetic  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main() -> None:
    torch.set_grad_enabled(False)
    from transformers.models.vibevoice_acoustic_tokenizer.configuration_vibevoice_acoustic_tokenizer import (
        VibeVoiceAcousticTokenizerConfig,
    )
    from transformers.models.vibevoice_acoustic_tokenizer.modeling_vibevoice_acoustic_tokenizer import (
        VibeVoiceAcousticTokenizerEncoderModel,
    )

    enc = VibeVoiceAcousticTokenizerEncoderModel(VibeVoiceAcousticTokenizerConfig().encoder_config)
    esd = synthetic.encoder_state_dict(2)
    print("encoder load:", enc.load_state_dict(esd, strict=True), sum(p.numel() for p in enc.parameters()))
    enc.eval()
    g = torch.Generator().manual_seed(99)
    audio = 0.3 * torch.randn(2, 1, 3 * 3200, generator=g)
    lat = enc(audio).latents
    print("latents", tuple(lat.shape), lat.abs().mean().item(), lat.abs().max().item())
    # causal prefix property: the first frame only needs the first 3200 samples
    lat1 = enc(audio[:, :, :3200]).latents
    print("prefix max diff", (lat1 - lat[:, :1]).abs().max().item())
    np.savez(os.path.join(OUT, "encoder_small.npz"), audio=audio.numpy(), latents=lat.numpy())


if __name__ == "__main__":
    main()

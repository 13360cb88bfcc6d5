"""Generate tests/golden/*.npz by running THE REFERENCE ITSELF (test infrastructure).

Runs only in the authoring container, where ``/root/reference`` is mounted:

    python oracle/make_golden.py            # writes tests/golden/*.npz

* DiT path: imports the reference's own PyTorch modules
  (``/root/reference/src/smalltts/models/backbone/{model,dit,style,phonemes}.py``) with
  import stubs for ``phonemizer``/``inflect`` (oracle/ref_stubs), loads the seeded weights
  of ``smalltts_b200.synthetic`` into ``DiTModel(64)`` with ``strict=True`` and records
  ``encode_conditions`` / ``denoise_step`` / the 4-step loop of infer/onnx.py:98-125.
* Vocoder: ``transformers`` 5.5.0 ``VibeVoiceAcousticTokenizerDecoderModel`` (the published
  restatement of the un-vendored ``decoder.onnx``), same treatment.

The fixtures hold inputs and outputs only; weights are re-drawn from the seed by the tests.
``/root/reference`` does not exist on the GPU box, so nothing else may import it.
"""

from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "ref_stubs"))
sys.path.insert(0, "/root/reference/src")

from smalltts_b200 import synthetic  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def ref_alpha_sigma(t: float):
    # verbatim call into the hot path's own schedule (infer/onnx.py:31-39; numpy, fp64 inside)
    from smalltts.infer.onnx import _get_alpha_sigma

    a, s = _get_alpha_sigma(t)
    return a, s


def main() -> None:
    torch.manual_seed(0)
    torch.set_grad_enabled(False)
    os.makedirs(OUT, exist_ok=True)

    from smalltts.models.backbone.model import DiTModel

    model = DiTModel(64)
    sd = synthetic.dit_state_dict(0)
    missing = model.load_state_dict(sd, strict=True)
    print("DiT load:", missing, sum(p.numel() for p in model.parameters()))
    model.eval()

    g = torch.Generator().manual_seed(1234)

    # ---- schedule (reference torch + its numpy twin is what the oracle restates)
    ts = np.linspace(1, 0, 4, dtype=np.float32)
    sched = np.array([ref_alpha_sigma(float(t)) for t in ts], dtype=np.float64)
    from smalltts.infer.onnx import _compute_rope_freqs

    np.savez(os.path.join(OUT, "schedule.npz"), t=ts, alpha_sigma=sched, rope=_compute_rope_freqs(9))

    # ---- ragged condition encode
    ref = torch.randn(2, 5, 64, generator=g)
    ref_len = torch.tensor([5, 3], dtype=torch.int64)
    ids = torch.randint(1, 198, (2, 7), generator=g)
    pmask = torch.tensor([[1] * 7, [1] * 4 + [0] * 3], dtype=torch.bool)
    ids = ids * pmask
    cached = model.encode_conditions(ref, ref_len, ids, pmask, 6)
    np.savez(
        os.path.join(OUT, "cond_small.npz"),
        ref=ref.numpy(), ref_len=ref_len.numpy(), ids=ids.numpy(), pmask=pmask.numpy(),
        ref_mask=cached["ref_mask"].numpy(),
        **{f"{k}_{i}": cached["layers"][i][k].numpy() for i in (0, 11) for k in ("k_ref", "v_ref", "k_text", "v_text")},
    )

    # ---- one ragged denoise step with per-utterance t
    x_t = torch.randn(2, 6, 64, generator=g)
    mask = torch.tensor([[1] * 6, [1] * 4 + [0] * 2], dtype=torch.bool)
    t = torch.tensor([0.7, 0.3])
    v = model.denoise_step(x_t, mask, t, cached)
    # full (uncached) forward must agree: model.py:57-86
    v_full = model(x_t, ref, ref_len, mask, ids, pmask, t)
    print("cached vs full max diff", (v - v_full).abs().max().item())
    np.savez(os.path.join(OUT, "denoise_small.npz"), x_t=x_t.numpy(), mask=mask.numpy(), t=t.numpy(), velocity=v.numpy())

    # ---- config 1 of BASELINE.json: B=1, 2 s (T=15), R=15, tokens 1..30 (bench.rs:22-23)
    refs, _, frames, noise = synthetic.synthetic_inputs(1, 15, 15, 30)
    ids1 = [list(range(1, 31))]
    ref1 = refs[0][None]
    cached1 = model.encode_conditions(ref1, torch.tensor([15]), torch.tensor(ids1), torch.ones(1, 30, dtype=torch.bool), 15)
    x_pred = torch.zeros(1, 15, 64)
    m1 = torch.ones(1, 15, dtype=torch.bool)
    for s_i, t_val in enumerate(ts):  # infer/onnx.py:102-125 with supplied noise
        a, s = ref_alpha_sigma(float(t_val))
        x_t1 = (float(a) * x_pred + float(s) * noise[s_i])
        vel = model.denoise_step(x_t1, m1, torch.tensor([float(t_val)]), cached1)
        x_pred = float(a) * x_t1 - float(s) * vel
    print("c1 latents", x_pred.abs().mean().item(), x_pred.abs().max().item())

    # ---- vocoder
    from transformers.models.vibevoice_acoustic_tokenizer.configuration_vibevoice_acoustic_tokenizer import (
        VibeVoiceAcousticTokenizerConfig,
    )
    from transformers.models.vibevoice_acoustic_tokenizer.modeling_vibevoice_acoustic_tokenizer import (
        VibeVoiceAcousticTokenizerDecoderModel,
    )

    dec = VibeVoiceAcousticTokenizerDecoderModel(VibeVoiceAcousticTokenizerConfig().decoder_config)
    vsd = synthetic.vocoder_state_dict(1)
    print("vocoder load:", dec.load_state_dict(vsd, strict=True), sum(p.numel() for p in dec.parameters()))
    dec.eval()

    lat = torch.randn(2, 3, 64, generator=g)
    audio = dec(lat.permute(0, 2, 1)).audio
    np.savez(os.path.join(OUT, "vocoder_small.npz"), latents=lat.numpy(), audio=audio.numpy())
    print("vocoder small", tuple(audio.shape), audio.abs().mean().item(), audio.abs().max().item())

    audio1 = dec(x_pred.permute(0, 2, 1)).audio
    np.savez(
        os.path.join(OUT, "e2e_c1.npz"),
        ref=ref1.numpy(), ids=np.array(ids1), noise=noise.numpy(), latents=x_pred.numpy(), audio=audio1[0].numpy(),
    )
    print("e2e c1", tuple(audio1.shape), audio1.abs().mean().item(), audio1.abs().max().item())


if __name__ == "__main__":
    main()

"""Generate tests/golden/teacher_small.npz with THE REFERENCE'S OWN modules (test infrastructure; authoring container only).

    python oracle/make_golden_teacher.py

BASELINE config 5 / SURVEY 8(a18).  The reference has no teacher inference script, so the fixture is built from the
pieces it does have: ``DiTModel.forward`` (models/backbone/model.py:57-86) evaluated on the 3-way classifier-free
guidance batch exactly as ``get_x_pred`` assembles it (scripts/train/dmd2/distill.py:74-103 -- that module itself
cannot be imported here: it needs accelerate and the discriminator/ASR/SV stacks) and the schedule
``get_alpha_sigma`` of train/utils.py:12-22, walked with the deterministic DDIM step those v-prediction identities
imply (train/utils.py:54-67; SURVEY 7 "No teacher sampler exists").
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


def main() -> None:
    torch.set_grad_enabled(False)
    from smalltts.models.backbone.model import DiTModel
    from smalltts.train.utils import get_alpha_sigma

    model = DiTModel(64)
    print("DiT load:", model.load_state_dict(synthetic.dit_state_dict(0), strict=True))
    model.eval()

    g = torch.Generator().manual_seed(4321)
    B, T, R, P, steps = 2, 6, 5, 7, 6
    ref = torch.randn(B, R, 64, generator=g)
    ref_len = torch.tensor([5, 3], dtype=torch.int64)
    ids = torch.randint(1, 198, (B, P), generator=g)
    pmask = torch.tensor([[1] * 7, [1] * 4 + [0] * 3], dtype=torch.bool)
    ids = ids * pmask
    mask = torch.tensor([[1] * 6, [1] * 4 + [0] * 2], dtype=torch.bool)
    noise = torch.randn(B, T, 64, generator=g)
    s_text, s_spk = 2.0, 1.5

    def cfg_velocity(x_t, t):  # distill.py:74-103, verbatim structure
        x3 = x_t.repeat(3, 1, 1)
        ref3 = torch.cat([ref, ref, torch.zeros_like(ref)], dim=0)
        len3 = torch.cat([ref_len, ref_len, torch.zeros_like(ref_len)], dim=0)
        ph3 = torch.cat([ids, torch.zeros_like(ids), ids], dim=0)
        pm3 = torch.cat([pmask, torch.zeros_like(pmask).to(dtype=torch.bool), pmask], dim=0)
        v3 = model(x3, ref3, len3, mask.repeat(3, 1), ph3, pm3, t.repeat(3))
        v_c, v_ut, v_us = v3.chunk(3, dim=0)
        return v_c + s_text * (v_c - v_ut) + s_spk * (v_c - v_us)

    ts = torch.linspace(1.0, 0.0, steps + 1)
    x = noise.clone()
    first_v = None
    for s in range(steps):
        t = torch.full((B,), float(ts[s]))
        v = cfg_velocity(x, t)
        if first_v is None:
            first_v = v.clone()
        a, sg = get_alpha_sigma(t)
        an, sn = get_alpha_sigma(torch.full((B,), float(ts[s + 1])))
        a, sg, an, sn = (z.view(-1, 1, 1) for z in (a, sg, an, sn))
        x0 = a * x - sg * v  # distill.py:127-130
        eps = sg * x + a * v
        x = an * x0 + sn * eps
    print("teacher latents", x.abs().mean().item(), x.abs().max().item(), "finite", bool(torch.isfinite(x).all()))
    os.makedirs(OUT, exist_ok=True)
    np.savez(os.path.join(OUT, "teacher_small.npz"), ref=ref.numpy(), ref_len=ref_len.numpy(), ids=ids.numpy(),
             pmask=pmask.numpy(), mask=mask.numpy(), noise=noise.numpy(), steps=np.int64(steps),
             cfg=np.array([s_text, s_spk], dtype=np.float32), first_velocity=first_v.numpy(), latents=x.numpy())


if __name__ == "__main__":
    main()

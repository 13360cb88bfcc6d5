"""Pin the CPU oracle against fixtures produced by the reference's own PyTorch modules
(oracle/make_golden.py).  fp32 on both sides; tolerances cover summation-order noise only."""
import os

import numpy as np
import torch

from conftest import GOLDEN
from oracle import smalltts_oracle as O


def _load(name):
    return {k: v for k, v in np.load(os.path.join(GOLDEN, name)).items()}


def _close(a, b, atol, rtol=1e-4):
    a = torch.as_tensor(a).float()
    b = torch.as_tensor(b).float()
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a - b).abs().max().item()
    assert torch.allclose(a, b, atol=atol, rtol=rtol), f"max abs err {err}"


def test_schedule_matches_reference():
    g = _load("schedule.npz")
    for t, (a, s) in zip(g["t"], g["alpha_sigma"]):
        oa, os_ = O.alpha_sigma(float(t))
        assert abs(float(oa) - a) < 1e-6 and abs(float(os_) - s) < 1e-6
    # SURVEY.md 8(a2) schedule values
    a, s = O.alpha_sigma(2 / 3)
    assert abs(a - 0.277350) < 1e-5 and abs(s - 0.960769) < 1e-5


def test_rope_table_matches_reference_formula():
    g = _load("schedule.npz")
    assert np.allclose(O.rope_angles(9).numpy(), g["rope"], rtol=1e-6, atol=1e-7)
    ang = O.rope_angles(5)
    assert ang.shape == (1, 5, 64)
    assert torch.equal(ang[0, :, 0], ang[0, :, 1])
    assert abs(ang[0, 3, 2].item() - 3 * 10000 ** (-2 / 64)) < 1e-6


@torch.inference_mode()
def test_encode_conditions_matches_reference(dit_sd):
    g = _load("cond_small.npz")
    cond = O.encode_conditions(dit_sd, torch.tensor(g["ref"]), torch.tensor(g["ref_len"]),
                               torch.tensor(g["ids"]), torch.tensor(g["pmask"]))
    assert np.array_equal(cond["ref_mask"].numpy(), g["ref_mask"])
    for i in (0, 11):
        for k in ("k_ref", "v_ref", "k_text", "v_text"):
            _close(cond["layers"][i][k], g[f"{k}_{i}"], atol=2e-5)


@torch.inference_mode()
def test_denoise_step_matches_reference(dit_sd):
    c = _load("cond_small.npz")
    g = _load("denoise_small.npz")
    cond = O.encode_conditions(dit_sd, torch.tensor(c["ref"]), torch.tensor(c["ref_len"]),
                               torch.tensor(c["ids"]), torch.tensor(c["pmask"]))
    v = O.denoise_step(dit_sd, torch.tensor(g["x_t"]), torch.tensor(g["mask"]), torch.tensor(g["t"]), cond)
    _close(v, g["velocity"], atol=5e-5)


@torch.inference_mode()
def test_vocoder_matches_reference(voc_sd):
    g = _load("vocoder_small.npz")
    audio = O.vocoder_decode(voc_sd, torch.tensor(g["latents"]))
    _close(audio, g["audio"], atol=2e-5)


@torch.inference_mode()
def test_config1_end_to_end_matches_reference(dit_sd, voc_sd):
    """BASELINE.json configs[0]: single 2 s utterance, batch 1 (T=15, R=15, tokens 1..30)."""
    g = _load("e2e_c1.npz")
    out = O.synthesize_batch(dit_sd, voc_sd, [torch.tensor(g["ref"][0])], [g["ids"][0].tolist()], [15],
                             torch.tensor(g["noise"]))
    assert out[0].shape == (1, 15 * 3200)
    _close(out[0], g["audio"], atol=1e-4)


@torch.inference_mode()
def test_ragged_row_equals_solo_run(dit_sd, voc_sd):
    """Masks isolate rows and the vocoder is causal: a short utterance inside a
    padded batch equals the same utterance run alone."""
    from smalltts_b200 import synthetic

    refs, ids, frames, noise = synthetic.synthetic_inputs(2, [5, 3], [4, 6], [9, 5], seed=7)
    both = O.synthesize_batch(dit_sd, voc_sd, refs, ids, frames, noise)
    solo = O.synthesize_batch(dit_sd, voc_sd, refs[1:], ids[1:], frames[1:], noise[:, 1:, :3])
    _close(both[1], solo[0], atol=1e-4)


@torch.inference_mode()
def test_teacher_sampler_matches_reference(dit_sd):
    """BASELINE config 5 / SURVEY 8(a18): 3-way CFG (distill.py:74-103) + DDIM walk, fixture built from the
    reference's DiTModel.forward and get_alpha_sigma by oracle/make_golden_teacher.py."""
    g = _load("teacher_small.npz")
    cond3 = O.cfg_conditions(dit_sd, torch.tensor(g["ref"]), torch.tensor(g["ref_len"]), torch.tensor(g["ids"]),
                             torch.tensor(g["pmask"]))
    s_text, s_spk = map(float, g["cfg"])
    mask, noise = torch.tensor(g["mask"]), torch.tensor(g["noise"])
    v = O.cfg_velocity(dit_sd, noise, mask, torch.ones(noise.shape[0]), cond3, s_text, s_spk)
    _close(v, g["first_velocity"], atol=2e-4)
    x = O.sample_teacher(dit_sd, cond3, mask, noise, int(g["steps"]), s_text, s_spk)
    valid = mask[..., None].expand_as(x).numpy()  # padded frames are never attended to nor decoded
    _close(x.numpy()[valid], g["latents"][valid], atol=1e-3)


@torch.inference_mode()
def test_codec_encoder_matches_reference():
    """codec encoder restatement vs transformers' VibeVoiceAcousticTokenizerEncoderModel
    (oracle/make_golden_encoder.py), plus the causal-prefix property."""
    from smalltts_b200 import synthetic

    esd = synthetic.encoder_state_dict(2)
    g = _load("encoder_small.npz")
    lat = O.codec_encode(esd, torch.tensor(g["audio"]))
    assert lat.shape == (2, 3, 64)
    _close(lat, g["latents"], atol=2e-5)
    lat1 = O.codec_encode(esd, torch.tensor(g["audio"][:, :, :3200 + 1234]))  # a ragged tail is floored away
    _close(lat1, g["latents"][:, :1], atol=2e-5)


def test_resample_hq_oracle_vs_reference_fixture():
    """infer/utils.py:7-23: the fixture was produced by the reference's own resample_hq (torchaudio, fp32 conv1d);
    the restatement accumulates in fp64, so the difference is torchaudio's fp32 summation error over ~4k taps."""
    g = np.load(os.path.join(GOLDEN, "resample_small.npz"))
    for sr in (44100, 48000, 16000, 22050, 8000, 32000):
        y = O.resample_hq(g[f"x_{sr}"], sr, 24000)
        assert y.shape == g[f"y_{sr}"].shape, sr
        assert float(np.abs(y - g[f"y_{sr}"]).max()) <= 5e-6, sr
    x = g["x_16000"]
    assert O.resample_hq(x, 24000, 24000) is not None and np.array_equal(O.resample_hq(x, 24000, 24000), x)


def test_resample_bank_shape_and_dc_gain():
    """Filter-bank facts the CUDA side relies on: K = 2*width + down taps per phase, `up` phases, and every phase
    sums to ~1 (unit DC gain: a constant signal stays constant away from the edges)."""
    for a, b in ((44100, 24000), (48000, 24000), (16000, 24000), (22050, 24000)):
        bank, width, down, up = O.resample_bank(a, b)
        assert bank.shape == (up, 2 * width + down) and bank.dtype == np.float32
        assert np.allclose(bank.astype(np.float64).sum(1), 1.0, atol=2e-3), (a, b)
    n = 30000
    y = O.resample_hq(np.ones((1, n), np.float32), 44100, 24000)
    mid = y[0, 4000:-4000]
    assert np.abs(mid - 1.0).max() < 2e-3

"""Operator- and path-level parity of the CUDA engine (through the C ABI) against
  (1) the committed fixtures produced by the reference's own PyTorch modules (tests/golden, oracle/make_golden.py),
  (2) the CPU oracle on seeded inputs, and
  (3) size-independent properties at BASELINE.json's full configuration (B=8 x 10 s).

Stated tolerance (the engine feeds bf16 operands with fp32 accumulation to the tensor cores; every norm, softmax,
RoPE, residual stream and the schedule are fp32):
    relative L2 error vs the fp32 reference  <= 2e-2   on latents, velocities, K/V caches and waveforms
    relative L2 error vs the oracle with bf16-rounded GEMM operands  <= 1e-2  (two independent bf16 rounding
        realisations differ by about as much as each differs from fp32; this only checks the error is of that kind)
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

TOL_FP32 = 2e-2
TOL_BF16_EMU = 1e-2
# The parity build (precision="tight": fp16 instead of bf16 GEMM / attention operands, csrc/op16.cuh; everything else --
# kernels, fp32 accumulation, fp32 norms / softmax / RoPE / residual streams -- identical).  fp16 carries 11 significand
# bits like TF32, so operand rounding is 8x finer than bf16's: measured 3.2e-4 on latents, 3.7e-4 on velocities and
# 6.4e-4 .. 6.8e-4 on waveforms (fast build: 3e-3 / 5.2e-3) and up to 1.8e-3 on the cross K/V caches behind the 12-layer
# style encoder, at the same speed.
TOL_TIGHT = 2.5e-3


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def _load(name):
    return {k: v for k, v in np.load(os.path.join(GOLDEN, name)).items()}


@pytest.fixture(scope="module")
def tts(dit_sd, voc_sd):
    from smalltts_b200.infer import SmallTTS

    t = SmallTTS(state_dicts=(dit_sd, voc_sd))
    yield t
    t.engine.close()


@pytest.fixture(scope="module")
def tts_tight(dit_sd, voc_sd):
    from smalltts_b200.infer import SmallTTS

    t = SmallTTS(state_dicts=(dit_sd, voc_sd), precision="tight")
    yield t
    t.engine.close()


def test_native_library_is_what_runs(tts):
    from smalltts_b200.engine import Engine

    maps = open("/proc/self/maps").read()
    assert "libsmalltts_b200.so" in maps
    before = Engine.launch_count()
    tts.engine.decode(np.zeros((1, 1, 64), dtype=np.float32))
    assert Engine.launch_count() > before


def test_encode_conditions_vs_reference_fixture(tts):
    g = _load("cond_small.npz")
    cond = tts.engine.encode_conditions(g["ref"], g["ref_len"], g["ids"], g["pmask"].sum(1))
    for i in (0, 11):
        for k in ("k_ref", "v_ref", "k_text", "v_text"):
            got, want = cond.read_kv(i, k), g[f"{k}_{i}"]
            valid = g["ref_mask"] if "ref" in k else g["pmask"]
            m = valid[:, None, :, None]
            assert rel_l2(got * m, want * m) <= TOL_FP32, (i, k, rel_l2(got * m, want * m))
    cond.free()


def test_denoise_step_vs_reference_fixture(tts):
    c, g = _load("cond_small.npz"), _load("denoise_small.npz")
    cond = tts.engine.encode_conditions(c["ref"], c["ref_len"], c["ids"], c["pmask"].sum(1))
    frames = g["mask"].sum(1)
    v = tts.engine.denoise_step(cond, g["x_t"], frames, g["t"])
    m = g["mask"][..., None]
    assert np.isfinite(v).all()
    assert rel_l2(v * m, g["velocity"] * m) <= TOL_FP32, rel_l2(v * m, g["velocity"] * m)
    # same call with device-resident tensors
    vd = tts.engine.denoise_step(cond, torch.from_numpy(g["x_t"]).cuda(), frames, g["t"]).cpu().numpy()
    assert np.array_equal(vd, v)
    cond.free()


def test_vocoder_vs_reference_fixture(tts):
    g = _load("vocoder_small.npz")
    audio = tts.engine.decode(g["latents"])
    assert audio.shape == (2, 3 * 3200)
    assert rel_l2(audio, g["audio"][:, 0]) <= TOL_FP32, rel_l2(audio, g["audio"][:, 0])


def test_config1_end_to_end_vs_reference_fixture(tts):
    """BASELINE.json configs[0]: single 2 s utterance, batch 1 -- latents and waveform of the reference loop."""
    g = _load("e2e_c1.npz")
    cond = tts.engine.encode_conditions(g["ref"], [15], g["ids"], [30])
    lat = tts.engine.sample(cond, [15], 15, noise=g["noise"])
    assert rel_l2(lat, g["latents"]) <= TOL_FP32, rel_l2(lat, g["latents"])
    cond.free()
    audio = tts.synthesize(g["ref"][0], g["ids"][0].tolist(), 2.0, noise=g["noise"])
    assert audio.shape == (1, 15 * 3200) and audio.dtype == np.float32
    assert rel_l2(audio, g["audio"]) <= TOL_FP32, rel_l2(audio, g["audio"])


def test_ragged_batch_vs_oracle_fp32_and_bf16_emulation(tts, dit_sd, voc_sd):
    from oracle import smalltts_oracle as O
    from smalltts_b200 import synthetic

    refs, ids, frames, noise = synthetic.synthetic_inputs(3, [9, 5, 12], [4, 9, 6], [11, 20, 5], seed=11)
    durs = [f * 3200 / 24000 + 1e-3 for f in frames]
    got = tts.synthesize_batch(refs, ids, durs, noise=noise.numpy())
    with torch.inference_mode():
        want = O.synthesize_batch(dit_sd, voc_sd, refs, ids, frames, noise)
        O.set_gemm_mode("bf16")
        try:
            emu = O.synthesize_batch(dit_sd, voc_sd, refs, ids, frames, noise)
        finally:
            O.set_gemm_mode("fp32")
    for i in range(3):
        assert got[i].shape == tuple(want[i].shape)
        assert rel_l2(got[i], want[i].numpy()) <= TOL_FP32, (i, rel_l2(got[i], want[i].numpy()))
        assert rel_l2(got[i], emu[i].numpy()) <= TOL_BF16_EMU, (i, rel_l2(got[i], emu[i].numpy()))


def test_full_size_config2_properties(tts):
    """B=8 x 10 s (T=75, R=15, P=120): rows are independent (a row equals its solo run), the vocoder is causal
    (decoding a prefix gives the prefix), outputs are finite and the on-device Philox path is reproducible."""
    from smalltts_b200 import synthetic

    refs, ids, frames, noise = synthetic.synthetic_inputs(8, 75, 15, 120)
    durs = [10.0] * 8
    full = tts.synthesize_batch(refs, ids, durs, noise=noise.numpy())
    assert all(a.shape == (1, 240000) and np.isfinite(a).all() for a in full)
    solo = tts.synthesize_batch(refs[3:4], ids[3:4], durs[3:4], noise=noise[:, 3:4].numpy())
    assert rel_l2(full[3], solo[0]) <= 2e-3, rel_l2(full[3], solo[0])
    lat = np.random.default_rng(0).standard_normal((2, 75, 64)).astype(np.float32)
    a_full = tts.engine.decode(lat)
    a_pre = tts.engine.decode(lat[:, :20])
    assert rel_l2(a_pre, a_full[:, : 20 * 3200]) <= 1e-5
    s1 = tts.synthesize_batch(refs[:2], ids[:2], durs[:2], seed=123)
    s2 = tts.synthesize_batch(refs[:2], ids[:2], durs[:2], seed=123)
    s3 = tts.synthesize_batch(refs[:2], ids[:2], durs[:2], seed=124)
    assert np.array_equal(s1[0], s2[0]) and not np.array_equal(s1[0], s3[0])


# ------------------------------------------------------------------ the configurations the bench numbers are quoted on
def _oracle_latents_and_audio(dit_sd, voc_sd, refs, ids, frames, noise, audio_rows):
    """CPU oracle at full size: latents of every row (DiT path, a few seconds) and the waveform of `audio_rows` only (the
    vocoder is 91 % of the oracle's time; rows are independent, so a subset pins the same arithmetic)."""
    from oracle import smalltts_oracle as O

    with torch.inference_mode():
        ref, ref_len, idt, pmask, mask = O.pad_batch(refs, ids, frames)
        cond = O.encode_conditions(dit_sd, ref, ref_len, idt, pmask)
        lat = O.sample(dit_sd, cond, mask, noise)
        audio = {i: O.vocoder_decode(voc_sd, lat[i : i + 1, : frames[i]])[0, 0].numpy() for i in audio_rows}
    return lat.numpy(), audio


@pytest.fixture(scope="module")
def config2_oracle(dit_sd, voc_sd):
    from smalltts_b200 import synthetic

    refs, ids, frames, noise = synthetic.synthetic_inputs(8, 75, 15, 120)
    want_lat, want_audio = _oracle_latents_and_audio(dit_sd, voc_sd, refs, ids, frames, noise, audio_rows=(0, 5))
    return refs, ids, frames, noise, want_lat, want_audio


def test_config2_headline_batch_vs_oracle(tts, config2_oracle):
    """BASELINE.json configs[1] EXACTLY as bench.py runs it (B=8, T=75, R=15, P=120, 4 DMD steps): the engine picks its
    full-size tiles, 148-CTA persistent schedules and second-wave paths here, so this is the numerical check of the
    configuration the headline number is quoted on.  Latents after the DMD loop for all 8 rows, waveforms for two."""
    from smalltts_b200.engine import pad_batch

    refs, ids, frames, noise, want_lat, want_audio = config2_oracle
    ref, ref_len, idt, ph_len = pad_batch(refs, ids, frames)
    cond = tts.engine.encode_conditions(ref, ref_len, idt, ph_len)
    lat = tts.engine.sample(cond, frames, 75, noise=noise.numpy())
    cond.free()
    for b in range(8):
        err = rel_l2(lat[b], want_lat[b])
        print("config2 latents row", b, "rel_l2", err)
        assert err <= TOL_FP32, (b, err)
    got = tts.synthesize_batch(refs, ids, [10.0] * 8, noise=noise.numpy())  # the fused plan / graph path of the bench
    again = tts.synthesize_batch(refs, ids, [10.0] * 8, noise=noise.numpy())  # second call replays the CUDA graphs
    for b, w in want_audio.items():
        assert got[b].shape == (1, 240000)
        err = rel_l2(got[b][0], w)
        print("config2 waveform row", b, "rel_l2", err)
        assert err <= TOL_FP32, (b, err)
        assert rel_l2(again[b][0], w) <= TOL_FP32
    # and with device-resident buffers (the `value` leg of bench.py)
    dev = tts.synthesize_batch(refs, ids, [10.0] * 8, noise=noise.numpy(), device_out=True)
    assert rel_l2(dev[5][0].cpu().numpy(), want_audio[5]) <= TOL_FP32


@pytest.mark.parametrize("mode", ["1", "2"])
def test_attention_inside_the_chained_kernel_vs_oracle(mode, dit_sd, voc_sd, config2_oracle, monkeypatch):
    """STTS_CHAIN_ATTN=1: a whole denoiser evaluation is ONE launch of the chained kernel (attention runs on tcgen05 as a
    phase between q|k|v|gate and to_out, csrc/chain_attn.cuh); =2: one launch per block.  Same bar as the default
    schedule on the headline batch, and far fewer kernels per step."""
    from smalltts_b200.engine import pad_batch
    from smalltts_b200.infer import SmallTTS

    refs, ids, frames, noise, want_lat, want_audio = config2_oracle
    monkeypatch.setenv("STTS_CHAIN_ATTN", mode)
    t = SmallTTS(state_dicts=(dit_sd, voc_sd))  # the schedule is read when the engine is created
    try:
        ref, ref_len, idt, ph_len = pad_batch(refs, ids, frames)
        cond = t.engine.encode_conditions(ref, ref_len, idt, ph_len)
        t.engine.sample(cond, frames, 75, noise=noise.numpy())  # the first call also builds the per-timestep adaLN tables
        n0 = t.engine.launch_count()
        lat = t.engine.sample(cond, frames, 75, noise=noise.numpy())
        launches = t.engine.launch_count() - n0
        cond.free()
        for b in range(8):
            err = rel_l2(lat[b], want_lat[b])
            print("chain attention mode", mode, "latents row", b, "rel_l2", err)
            assert err <= TOL_FP32, (b, err)
        # per evaluation: 4 input-embedding GEMMs + stats/cast + 1 (or 13) chained launches, against 25 by default
        print("kernels in the 4-step DMD loop:", launches)
        assert launches <= (4 * 8 + 10 if mode == "1" else 4 * 20 + 10)
        got = t.synthesize_batch(refs, ids, [10.0] * 8, noise=noise.numpy())
        for b, w in want_audio.items():
            assert rel_l2(got[b][0], w) <= TOL_FP32
    finally:
        t.engine.close()


def test_tight_precision_mode_vs_reference_fixtures_and_oracle(tts_tight, tts, config2_oracle):
    """precision="tight" (fp16 operands, 11-bit significand like TF32): the reference's fp32 results within TOL_TIGHT on
    the reference-made fixtures (K/V caches, velocity, vocoder, config 1 end to end) and on the headline batch (B=8 x
    10 s) against the oracle -- an order of magnitude inside the fast build's tolerance, and closer than the fast build
    on every one of them."""
    from smalltts_b200.engine import pad_batch

    eng = tts_tight.engine
    assert eng.precision == "tight"
    c, g = _load("cond_small.npz"), _load("denoise_small.npz")
    cond = eng.encode_conditions(c["ref"], c["ref_len"], c["ids"], c["pmask"].sum(1))
    for i in (0, 11):
        for k in ("k_ref", "v_ref", "k_text", "v_text"):
            valid = (c["ref_mask"] if "ref" in k else c["pmask"])[:, None, :, None]
            err = rel_l2(cond.read_kv(i, k) * valid, c[f"{k}_{i}"] * valid)
            print("tight cache", i, k, err)
            assert err <= TOL_TIGHT, (i, k, err)
    m = g["mask"][..., None]
    v = eng.denoise_step(cond, g["x_t"], g["mask"].sum(1), g["t"])  # per-utterance t: the generic path
    err = rel_l2(v * m, g["velocity"] * m)
    print("tight velocity", err)
    assert err <= TOL_TIGHT
    cond.free()
    voc = _load("vocoder_small.npz")
    err = rel_l2(eng.decode(voc["latents"]), voc["audio"][:, 0])
    print("tight vocoder", err)
    assert err <= TOL_TIGHT
    e1 = _load("e2e_c1.npz")
    audio = tts_tight.synthesize(e1["ref"][0], e1["ids"][0].tolist(), 2.0, noise=e1["noise"])
    err, err_fast = rel_l2(audio, e1["audio"]), rel_l2(tts.synthesize(e1["ref"][0], e1["ids"][0].tolist(), 2.0, noise=e1["noise"]), e1["audio"])
    print("tight config1 waveform", err, "fast", err_fast)
    assert err <= TOL_TIGHT and err < err_fast
    # headline batch: chained DiT kernels, full-size vocoder tiles
    refs, ids, frames, noise, want_lat, want_audio = config2_oracle
    ref, ref_len, idt, ph_len = pad_batch(refs, ids, frames)
    cond = eng.encode_conditions(ref, ref_len, idt, ph_len)
    lat = eng.sample(cond, frames, 75, noise=noise.numpy())
    cond.free()
    worst = max(rel_l2(lat[b], want_lat[b]) for b in range(8))
    print("tight config2 latents (worst row)", worst)
    assert worst <= TOL_TIGHT
    got = tts_tight.synthesize_batch(refs, ids, [10.0] * 8, noise=noise.numpy())
    for b, w in want_audio.items():
        err = rel_l2(got[b][0], w)
        print("tight config2 waveform row", b, err)
        assert err <= TOL_TIGHT


def test_config3_clone_16_prompts_vs_oracle(tts, dit_sd, voc_sd):
    """BASELINE.json configs[2]: one 3 s reference (R=22) shared by 16 prompts with T ~ U{15..75}, P = round(1.53 T)
     in one ragged engine call; every row's latents and three waveforms against the oracle."""
    from smalltts_b200 import synthetic
    from smalltts_b200.engine import pad_batch

    rng = np.random.default_rng(3)
    frames = [int(x) for x in rng.integers(15, 76, size=16)]
    frames[0], frames[1] = 75, 15
    phon = [int(round(1.53 * f)) for f in frames]
    refs, ids, frames, noise = synthetic.synthetic_inputs(16, frames, 22, phon, seed=303)
    refs = [refs[0]] * 16  # one cloned voice
    want_lat, want_audio = _oracle_latents_and_audio(dit_sd, voc_sd, refs, ids, frames, noise, audio_rows=(0, 1, 9))
    ref, ref_len, idt, ph_len = pad_batch(refs, ids, frames)
    cond = tts.engine.encode_conditions(ref, ref_len, idt, ph_len)
    lat = tts.engine.sample(cond, frames, max(frames), noise=noise.numpy())
    cond.free()
    for b in range(16):
        err = rel_l2(lat[b, : frames[b]], want_lat[b, : frames[b]])
        assert err <= TOL_FP32, (b, frames[b], err)
    durs = [f * 3200 / 24000 + 1e-3 for f in frames]
    got = tts.synthesize_batch(refs, ids, durs, noise=noise.numpy())
    for b, w in want_audio.items():
        assert got[b].shape == (1, frames[b] * 3200)
        err = rel_l2(got[b][0], w)
        print("config3 waveform row", b, "frames", frames[b], "rel_l2", err)
        assert err <= TOL_FP32, (b, err)


def test_config4_mixed_ragged_slice_vs_oracle(tts, dit_sd, voc_sd):
    """BASELINE.json configs[3]-shaped slice: 16 of the 64 mixed 2-10 s prompts as one GPU's share (T ~ U{15..75},
    R ~ U{8..64} like data/dummy.py:32, P = round(1.53 T)), run the way parallel.py runs a shard: length-bucketed
    micro-batches.  Every row's waveform prefix (first 2 s) and full latents against the oracle."""
    from smalltts_b200 import parallel, synthetic
    from smalltts_b200.engine import pad_batch

    rng = np.random.default_rng(4)
    frames = [int(x) for x in rng.integers(15, 76, size=16)]
    rlen = [int(x) for x in rng.integers(8, 65, size=16)]
    phon = [int(round(1.53 * f)) for f in frames]
    refs, ids, frames, noise = synthetic.synthetic_inputs(16, frames, rlen, phon, seed=404)
    want_lat, want_audio = _oracle_latents_and_audio(dit_sd, voc_sd, refs, ids, frames, noise, audio_rows=(2, 11))
    ref, ref_len, idt, ph_len = pad_batch(refs, ids, frames)
    cond = tts.engine.encode_conditions(ref, ref_len, idt, ph_len)
    lat = tts.engine.sample(cond, frames, max(frames), noise=noise.numpy())
    cond.free()
    for b in range(16):
        err = rel_l2(lat[b, : frames[b]], want_lat[b, : frames[b]])
        assert err <= TOL_FP32, (b, frames[b], err)
    # the shard runner: micro-batches of similar length, results back in request order
    durs = [f * 3200 / 24000 + 1e-3 for f in frames]
    order = sorted(range(16), key=lambda i: (-frames[i], i))
    batches = parallel.length_buckets(order, frames, max_batch=8)
    assert sorted(i for mb in batches for i in mb) == list(range(16))
    out = [None] * 16
    for mb in batches:
        tl = max(frames[i] for i in mb)
        res = tts.synthesize_batch([refs[i] for i in mb], [ids[i] for i in mb], [durs[i] for i in mb],
                                   noise=np.ascontiguousarray(noise.numpy()[:, mb, :tl]))
        for i, a in zip(mb, res):
            out[i] = a
    for b, w in want_audio.items():
        err = rel_l2(out[b][0], w)
        print("config4 waveform row", b, "frames", frames[b], "rel_l2", err)
        assert err <= TOL_FP32, (b, err)


def test_philox_noise_is_standard_normal(tts):
    """sample() with on-device noise at alpha~0 (one step at t=1) returns x_pred = a*x_t - s*v; check only the
    generator through a 1-step run with zeroed velocity weights is not possible here, so test moments of
    x_t indirectly: two seeds differ and outputs stay finite.  Moments are checked in test_reference_api."""
    from smalltts_b200 import synthetic

    refs, ids, frames, _ = synthetic.synthetic_inputs(1, 8, 4, 6)
    a = tts.synthesize_batch(refs, ids, [8 * 3200 / 24000 + 1e-3], seed=5)[0]
    assert np.isfinite(a).all() and a.std() > 0


def test_reference_api_surface(tts):
    from smalltts_b200 import infer

    assert (infer.SAMPLE_RATE, infer.HOP_SIZE, infer.NUM_STEPS, infer.CHARS_PER_SECOND) == (24000, 3200, 4, 11.5)
    assert infer.estimate_duration("x" * 23) == 2.0 and infer.estimate_duration("") == 0.5
    conds = [torch.randn(6, 64), torch.randn(9, 64)]
    out = tts(conds, [[1, 2, 3], [4, 5]], [[6, 7], [8, 9, 10]], duration_sec=1.0)
    assert len(out) == 2 and all(isinstance(o, torch.Tensor) and o.shape == (1, 7 * 3200) for o in out)
    with pytest.raises(ValueError):
        tts.synthesize_batch([np.zeros((4, 63), np.float32)], [[1]], [1.0])


def test_teacher_sampler_vs_reference_fixture(tts):
    """BASELINE config 5: 3-way CFG + DDIM through stts_sample_teacher vs the fixture produced by
    the reference's DiTModel.forward (oracle/make_golden_teacher.py)."""
    g = _load("teacher_small.npz")
    B, T = g["noise"].shape[:2]
    frames = g["mask"].sum(1).tolist()
    cond3 = tts.engine.encode_conditions_cfg(g["ref"], g["ref_len"], g["ids"], g["pmask"].sum(1))
    assert cond3.B == 3 * B
    s_text, s_spk = map(float, g["cfg"])
    x = tts.engine.sample_teacher(cond3, frames, T, steps=int(g["steps"]), cfg_text=s_text, cfg_speaker=s_spk,
                                  noise=g["noise"])
    cond3.free()
    valid = np.broadcast_to(g["mask"][..., None], x.shape)
    err = rel_l2(x[valid], g["latents"][valid])
    print("teacher latents rel_l2", err)
    assert np.isfinite(x).all() and err <= TOL_FP32


def test_teacher_sampler_guidance_identity(tts):
    """With both guidance scales at 0 the CFG velocity is the conditional one, so a 1-step teacher walk from
    x_1 = noise must equal the first DMD step (x_pred = alpha x_t - sigma v at t = 1 up to sigma'(0) = 3e-5)."""
    from smalltts_b200 import synthetic

    refs, ids, frames, noise = synthetic.synthetic_inputs(2, [9, 7], [6, 4], [11, 8], seed=11)
    from smalltts_b200.engine import pad_batch

    ref, ref_len, idt, ph_len = pad_batch(refs, ids, frames)
    T = max(frames)
    cond3 = tts.engine.encode_conditions_cfg(ref, ref_len, idt, ph_len)
    x = tts.engine.sample_teacher(cond3, frames, T, steps=1, cfg_text=0.0, cfg_speaker=0.0, noise=noise[0].numpy())
    cond3.free()
    cond = tts.engine.encode_conditions(ref, ref_len, idt, ph_len)
    n4 = np.zeros((1,) + tuple(noise[0].shape), dtype=np.float32)
    n4[0] = noise[0].numpy()
    y = tts.engine.sample(cond, frames, T, noise=n4, steps=1, timesteps=[1.0])
    cond.free()
    for b in range(2):
        assert rel_l2(x[b, : frames[b]], y[b, : frames[b]]) <= 2e-3


@pytest.fixture(scope="module")
def tts_enc(dit_sd, voc_sd):
    """Engine that also carries the codec encoder (clone path, BASELINE config 3)."""
    from smalltts_b200 import synthetic
    from smalltts_b200.infer import SmallTTS

    t = SmallTTS(state_dicts=(dit_sd, voc_sd, synthetic.encoder_state_dict(2)))
    yield t
    t.engine.close()


def test_codec_encoder_vs_reference_fixture(tts_enc):
    """stts_encode_audio vs transformers' VibeVoiceAcousticTokenizerEncoderModel fixture."""
    g = _load("encoder_small.npz")
    lat = tts_enc.engine.encode_audio(g["audio"])
    assert lat.shape == g["latents"].shape
    err = rel_l2(lat, g["latents"])
    print("encoder latents rel_l2", err)
    assert np.isfinite(lat).all() and err <= TOL_FP32


def test_codec_encoder_vs_oracle_ragged_and_causal(tts_enc):
    """3 s reference (clone.py): a tail shorter than one hop is floored away, and a prefix of the audio gives a prefix
    of the latents (causal convolutions), checked against the CPU oracle on the full clip."""
    import torch

    from oracle import smalltts_oracle as O
    from smalltts_b200 import synthetic

    g = torch.Generator().manual_seed(5)
    audio = 0.3 * torch.randn(1, 1, 7 * 3200 + 777, generator=g)
    lat = tts_enc.engine.encode_audio(audio.numpy())
    assert lat.shape == (1, 7, 64)
    with torch.inference_mode():
        want = O.codec_encode(synthetic.encoder_state_dict(2), audio).numpy()
    assert rel_l2(lat, want) <= TOL_FP32
    head = tts_enc.engine.encode_audio(audio.numpy()[:, :, : 4 * 3200])
    assert rel_l2(head, lat[:, :4]) <= 1e-5


def test_clone_path_config3(tts_enc):
    """BASELINE config 3: one reference wav -> latents on the engine's encoder, shared by several prompts in one
    batched call; each row must equal the same prompt synthesised alone with that voice."""
    import torch

    from smalltts_b200 import synthetic

    g = torch.Generator().manual_seed(6)
    wav = 0.2 * torch.randn(3 * 24000, generator=g).numpy()
    ref = tts_enc.clone_voice(wav)
    assert ref.shape == (22, 64)  # 3 s -> 22 whole hops
    _, ids, frames, noise = synthetic.synthetic_inputs(3, [9, 6, 12], 1, [14, 9, 18], seed=21)
    durs = [f * 3200 / 24000 + 1e-3 for f in frames]
    both = tts_enc.synthesize_batch([ref] * 3, ids, durs, noise=noise.numpy())
    solo = tts_enc.synthesize_batch([ref], ids[1:2], durs[1:2], noise=noise.numpy()[:, 1:2, : frames[1]])
    assert both[1].shape == (1, frames[1] * 3200)
    assert rel_l2(both[1], solo[0]) <= 2e-3


def test_edge_cases_single_frame_empty_text_and_long_utterance(tts, dit_sd, voc_sd):
    """Edge cases of the path: a 1-frame utterance, a row without any phoneme (all keys of that source masked), a row
    without reference frames, and a 30 s utterance (the cap of estimate_duration, infer/onnx.py:17-18) in one ragged
    batch -- every row against the CPU oracle."""
    import torch

    from oracle import smalltts_oracle as O
    from smalltts_b200 import synthetic

    frames = [1, 7, 225, 12]
    refs, ids, _, noise = synthetic.synthetic_inputs(4, frames, [3, 1, 15, 6], [2, 1, 40, 9], seed=31)
    ids[1] = []          # no text at all
    refs[3] = refs[3][:0]  # no reference audio at all
    durs = [f * 3200 / 24000 + 1e-3 for f in frames]
    got = tts.synthesize_batch(refs, ids, durs, noise=noise.numpy())
    with torch.inference_mode():
        want = O.synthesize_batch(dit_sd, voc_sd, refs, ids, frames, noise)
    for i, (g, w) in enumerate(zip(got, want)):
        assert g.shape == (1, frames[i] * 3200) and np.isfinite(g).all()
        err = rel_l2(g, w.numpy())
        print("edge row", i, "frames", frames[i], "rel_l2", err)
        assert err <= TOL_FP32


def test_invalid_arguments_raise(tts):
    """Error behaviour of the boundary: bad lengths / shapes come back as ValueError (status -1), never a crash."""
    from smalltts_b200 import synthetic

    refs, ids, frames, _ = synthetic.synthetic_inputs(1, 4, 3, 5, seed=1)
    with pytest.raises(ValueError):
        tts.synthesize_batch(refs, ids, [])  # empty batch
    with pytest.raises(ValueError):
        tts.engine.synthesize(np.zeros((1, 3, 64), np.float32), [4], np.zeros((1, 5), np.int64), [5], [4], 4)  # ref_len > R
    with pytest.raises(ValueError):
        tts.engine.synthesize(np.zeros((1, 3, 64), np.float32), [3], np.zeros((1, 5), np.int64), [5], [0], 4)  # 0 frames
    with pytest.raises((ValueError, RuntimeError)):
        tts.engine.encode_audio(np.zeros((1, 6400), np.float32))  # this engine carries no encoder weights


# ------------------------------------------------------------------ clone path: resampler + standalone codec classes
TOL_RESAMPLE = 2e-5  # abs, unit-amplitude signals: fp32 accumulation over ~4k filter taps (torchaudio itself: 2e-6 vs fp64)


def test_resample_hq_vs_reference_fixture(tts):
    """stts_resample vs the reference's own resample_hq outputs (tests/golden/resample_small.npz, made by
    oracle/make_golden_resample.py) at the common wav rates; host and device buffers."""
    import torch

    g = _load("resample_small.npz")
    for sr in (44100, 48000, 16000, 22050, 8000, 32000):
        x, want = g[f"x_{sr}"], g[f"y_{sr}"]
        y = tts.engine.resample(x, sr, 24000)
        assert y.shape == want.shape and y.dtype == np.float32
        err = float(np.abs(y - want).max())
        print("resample", sr, "max abs err", err)
        assert err <= TOL_RESAMPLE, sr
        yd = tts.engine.resample(torch.from_numpy(x).cuda(), sr, 24000)
        assert yd.is_cuda and float(np.abs(yd.cpu().numpy() - y).max()) <= 1e-7
    same = tts.engine.resample(g["x_8000"], 24000, 24000)
    assert np.array_equal(same, g["x_8000"])  # infer/utils.py:20-21


def test_resample_hq_long_clip_vs_oracle_and_errors(tts):
    """A 3 s clip at 44.1 kHz (clone.py's typical input; 80 phases x 4151 taps, several CTAs per phase group) against
    the fp64 oracle; unsupported rate pairs fail loudly instead of allocating a gigantic filter bank."""
    import torch

    from oracle import smalltts_oracle as O

    g = torch.Generator().manual_seed(77)
    x = (0.3 * torch.randn(1, 3 * 44100 + 13, generator=g)).numpy()
    y = tts.engine.resample(x, 44100, 24000)
    want = O.resample_hq(x, 44100, 24000)
    assert y.shape == want.shape == (1, 72008)
    assert float(np.abs(y - want).max()) <= TOL_RESAMPLE
    with pytest.raises(ValueError):
        tts.engine.resample(x, 44101, 24000)  # coprime rates: a 24000-phase bank
    with pytest.raises(ValueError):
        tts.engine.resample(x[0], 44100, 24000)  # not (B, N)


def test_clone_voice_resamples_on_device(tts_enc):
    """clone.py:27-36 with a 44.1 kHz stereo wav: down-mix, resample_hq, encode -- all on the engine -- equals the
    encoder applied to the oracle's resampled audio."""
    import torch

    from oracle import smalltts_oracle as O

    g = torch.Generator().manual_seed(8)
    wav = (0.2 * torch.randn(2, 2 * 44100, generator=g)).numpy()
    ref = tts_enc.clone_voice(wav, sample_rate=44100)
    mono24 = O.resample_hq(wav.mean(0, keepdims=True), 44100, 24000)
    want = tts_enc.engine.encode_audio(mono24)[0]
    assert ref.shape == want.shape == (15, 64)
    # the two waveforms differ by ~1e-6 (fp32 vs fp64 FIR accumulation); the encoder's bf16 operand rounding turns that
    # into flips of the last bf16 bit, i.e. the same kind of error as two bf16 realisations (measured 3.5e-3)
    assert rel_l2(ref, want) <= TOL_BF16_EMU

    from smalltts_b200.utils import resample_hq

    y = resample_hq(torch.from_numpy(wav), 44100, 24000, engine=tts_enc.engine)
    assert tuple(y.shape) == (2, 48000) and y.dtype == torch.float32
    assert resample_hq(y, 24000, 24000) is y


def test_standalone_codec_classes_run_on_codec_only_engines(tts, tts_enc, voc_sd):
    """codec/onnx.py:34-75: Decoder / Encoder are usable without a DiT (sv.py:24).  A codec-only engine gives the
    same audio as the full engine and refuses the DiT operators loudly."""
    import torch

    from smalltts_b200 import synthetic
    from smalltts_b200.codec import Decoder, Encoder

    g = torch.Generator().manual_seed(9)
    lat = torch.randn(2, 4, 64, generator=g)
    dec = Decoder(state_dict=voc_sd)
    audio = dec.decode(lat)
    assert tuple(audio.shape) == (2, 1, 4 * 3200)
    assert rel_l2(audio[:, 0].numpy(), tts.engine.decode(lat.numpy())) <= 1e-6
    with pytest.raises(RuntimeError, match="not loaded"):
        dec.engine.encode_conditions(np.zeros((1, 2, 64), np.float32), [2], np.ones((1, 3), np.int64), [3])
    with pytest.raises(RuntimeError, match="not loaded"):
        dec.engine.encode_audio(np.zeros((1, 3200), np.float32))
    dec.engine.close()

    enc = Encoder(state_dict=synthetic.encoder_state_dict(2))
    wav = 0.3 * torch.randn(1, 1, 2 * 3200, generator=g)
    lat2 = enc.encode(wav)
    assert tuple(lat2.shape) == (1, 2, 64)
    assert rel_l2(lat2.numpy(), tts_enc.engine.encode_audio(wav.numpy())) <= 1e-6
    with pytest.raises(RuntimeError, match="not loaded"):
        enc.engine.decode(lat.numpy())
    enc.engine.close()
    shared = Decoder(engine=tts.engine)
    assert rel_l2(shared.decode(lat.cuda()).cpu().numpy(), audio.numpy()) <= 1e-6


def test_serving_pipeline_and_microbatcher(tts_enc):
    """pipeline.rs:60-112 on the engine: synthesize_timed returns ceil(duration*7.5) frames of audio and the stage
    timings; requests submitted together share one engine pass and each equals its solo result (same seed)."""
    import torch

    from smalltts_b200 import serve

    g = torch.Generator().manual_seed(31)
    refs = [(0.2 * torch.randn(n, generator=g)).numpy() for n in (2 * 24000, 24000 + 500, 3 * 24000)]
    toks = [[5, 9, 20, 33], [7, 7, 12], [101, 3, 44, 9, 2]]
    durs = [1.01, 0.5, 2.0]
    pipe = serve.Pipeline(tts_enc, shape_buckets=None)  # exact shapes here; bucketing has its own test below
    tts_enc._seed, tts_enc._calls = 123, 0
    audio, tm = pipe.synthesize_timed(refs[0], toks[0], durs[0])
    assert audio.shape == (8 * 3200,) and np.isfinite(audio).all()  # ceil(1.01 * 7.5) = 8 frames
    assert tm.codec_enc_ms > 0 and tm.cond_enc_ms > 0 and tm.denoise_ms > 0 and tm.codec_dec_ms > 0
    assert tm.total_ms >= tm.codec_enc_ms + tm.cond_enc_ms + tm.denoise_ms + tm.codec_dec_ms - 1e-3

    tts_enc._calls = 0
    many, tm3 = pipe.synthesize_many(refs, toks, durs)
    assert [a.shape[0] for a in many] == [8 * 3200, 4 * 3200, 15 * 3200] and tm3.batch == 3

    b = serve.MicroBatcher(pipe.synthesize_many, max_batch=4, max_wait_ms=500)
    try:
        tts_enc._calls = 0
        futs = [b.submit(r, t, d) for r, t, d in zip(refs, toks, durs)]
        res = [f.result(120) for f in futs]
    finally:
        b.close()
    assert b.batches_run == 1 and res[0][1].batch == 3
    for (a, _), want in zip(res, many):
        assert rel_l2(a, want) <= 1e-6  # same seed, same batch composition -> same pass


def test_devices_kwarg_splits_the_batch_over_gpus(tts, dit_sd, voc_sd):
    """SmallTTS(devices=[0, 1]) (replicas + batch split, no data-path collective): with supplied noise
    the waveforms equal the single-GPU ones.  Needs two visible GPUs."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 visible GPUs (runs under gpurun --gpus 2)")
    from smalltts_b200 import synthetic
    from smalltts_b200.infer import SmallTTS

    refs, ids, frames, noise = synthetic.synthetic_inputs(5, [12, 5, 9, 12, 7], [4, 6, 3, 5, 4], [10, 8, 12, 6, 9], seed=5)
    durs = [f * 3200 / 24000 + 1e-3 for f in frames]
    want = tts.synthesize_batch(refs, ids, durs, noise=noise.numpy())
    multi = SmallTTS(state_dicts=(dit_sd, voc_sd), devices=[0, 1])
    try:
        got = multi.synthesize_batch(refs, ids, durs, noise=noise.numpy())
    finally:
        for r in multi._replicas:
            r.engine.close()
        multi.engine.close()
    assert [g.shape for g in got] == [w.shape for w in want]
    for g, w in zip(got, want):
        assert rel_l2(g, w) <= 2e-3  # batch composition changes tile shapes, not the arithmetic per row


def test_engine_clone_shares_weights_and_runs_concurrently(tts):
    """stts_engine_clone: same weights, own streams / plans.  Same inputs + same noise -> same audio as the source; two
    handles driven from two host threads at once give the same results as one after the other."""
    import threading

    from smalltts_b200 import synthetic
    from smalltts_b200.engine import pad_batch

    refs, ids, frames, noise = synthetic.synthetic_inputs(3, [9, 6, 12], [4, 6, 3], [14, 9, 18], seed=33)
    ref, ref_len, idt, ph_len = pad_batch(refs, ids, frames)
    want = tts.engine.synthesize(ref, ref_len, idt, ph_len, frames, 12, noise=noise.numpy())
    clone = tts.engine.clone()
    try:
        got = clone.synthesize(ref, ref_len, idt, ph_len, frames, 12, noise=noise.numpy())
        assert rel_l2(got, want) <= 1e-6
        with pytest.raises(RuntimeError, match="clone"):
            clone.load_state_dicts({"x": np.zeros(1, np.float32)}, None)
        outs = {}

        def work(name, eng):
            for _ in range(5):
                outs[name] = eng.synthesize(ref, ref_len, idt, ph_len, frames, 12, noise=noise.numpy())

        ths = [threading.Thread(target=work, args=(n, e)) for n, e in (("a", tts.engine), ("b", clone))]
        [t.start() for t in ths]
        [t.join() for t in ths]
        assert rel_l2(outs["a"], want) <= 1e-6 and rel_l2(outs["b"], want) <= 1e-6
    finally:
        clone.close()


def test_shape_buckets_do_not_change_the_audio(dit_sd, voc_sd):
    """Rounding the padded (R, P, T) up only adds masked rows: with supplied noise (which pins T) the audio equals the
    unbucketed run; with the on-device stream every utterance still gets its own length."""
    from smalltts_b200 import synthetic
    from smalltts_b200.infer import SmallTTS

    refs, ids, frames, noise = synthetic.synthetic_inputs(3, [9, 6, 12], [4, 6, 3], [14, 9, 18], seed=41)
    durs = [f * 3200 / 24000 + 1e-3 for f in frames]
    plain = SmallTTS(state_dicts=(dit_sd, voc_sd))
    bucketed = SmallTTS(state_dicts=(dit_sd, voc_sd), shape_buckets=(8, 16, 5))
    try:
        a = plain.synthesize_batch(refs, ids, durs, noise=noise.numpy())
        b = bucketed.synthesize_batch(refs, ids, durs, noise=noise.numpy())
        for x, y in zip(a, b):
            assert rel_l2(y, x) <= 2e-3
        c = bucketed.synthesize_batch(refs, ids, durs, seed=5)  # T 12 -> 15: padded beyond every utterance
        assert [x.shape for x in c] == [(1, f * 3200) for f in frames] and all(np.isfinite(x).all() for x in c)
    finally:
        plain.engine.close()
        bucketed.engine.close()


def test_files_to_engine_to_audio(tts_enc, dit_sd, voc_sd, tmp_path):
    """SURVEY 8(f2) on hardware, the constructor path of infer/onnx.py:53-66: SmallTTS built from weight FILES -- the
    DiT as a packed .sttsw container, the codec decoder as .safetensors, the codec encoder as a trainer-style .pt
    checkpoint with EMA / DDP / torch.compile prefixes (distill.py:39-57) -- must give bit-identical audio and
    reference latents to the engine that was handed the same tensors as state dicts."""
    from safetensors.numpy import save_file

    from smalltts_b200 import synthetic, weights
    from smalltts_b200.infer import SmallTTS

    enc_sd = synthetic.encoder_state_dict(2)
    dit_path, voc_path, enc_path = tmp_path / "dit.sttsw", tmp_path / "decoder.safetensors", tmp_path / "encoder.pt"
    weights.save_packed(str(dit_path), dit_sd)
    save_file({k: np.ascontiguousarray(v.numpy()) for k, v in voc_sd.items()}, str(voc_path))
    wrapped = {"ema_model.module._orig_mod." + k: v for k, v in enc_sd.items()}
    wrapped.update({"initted": torch.tensor(True), "step": torch.tensor(7)})
    torch.save({"student_model": wrapped}, enc_path)

    from_files = SmallTTS(str(dit_path), None, str(voc_path), codec_encoder_path=str(enc_path))
    try:
        refs, ids, frames, noise = synthetic.synthetic_inputs(2, [9, 14], [5, 8], [12, 20], seed=77)
        durs = [f * 3200 / 24000 + 1e-3 for f in frames]
        want = tts_enc.synthesize_batch(refs, ids, durs, noise=noise.numpy())
        got = from_files.synthesize_batch(refs, ids, durs, noise=noise.numpy())
        for g, w in zip(got, want):
            assert g.shape == w.shape and np.array_equal(g, w)
        wav = (0.2 * torch.randn(2 * 24000, generator=torch.Generator().manual_seed(3))).numpy()
        assert np.array_equal(from_files.clone_voice(wav), tts_enc.clone_voice(wav))
    finally:
        from_files.engine.close()
    with pytest.raises((KeyError, ValueError, RuntimeError)):
        SmallTTS(str(voc_path), None, str(voc_path))  # a decoder file is not a DiT: rejected by name, before any upload

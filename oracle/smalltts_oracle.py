"""CPU oracle for the smalltts hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain fp32 torch-CPU restatement of the arithmetic behind
``SmallTTS.synthesize`` in the reference (condition encoder -> 4-step DMD
denoiser loop -> VibeVoice codec decoder).  It exists so that ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs can check (and time) the CUDA engine against it.  Nothing
under ``smalltts_b200/`` may import it: the product path has no CPU fallback.

Every function is stateless and driven by a flat ``dict[str, Tensor]`` that
uses the reference's own ``state_dict`` key names, so
real checkpoints drop in unchanged.

Pinning status
--------------
* DiT / condition path: the reference ships no golden vectors or numeric
  tests for this path.  The restatement is pinned against
  OUTPUTS OF THE REFERENCE ITSELF: ``oracle/make_golden.py`` imports the
  reference's PyTorch modules from ``/root/reference`` in the authoring
  container, loads the seeded weights of ``oracle/weights.py`` into them and
  stores input/output fixtures under ``tests/golden/``;
  ``tests/test_oracle_golden.py`` replays them through this file.
* Vocoder: the arithmetic lives in a third-party artefact that is not under
  ``/root/reference`` (``assets/codec/decoder.onnx`` of the HF repo
  smallbraineng/smalltts, an export of microsoft/VibeVoice's acoustic
  tokenizer decoder; executed by onnxruntime 1.22.1 per the reference's
  uv.lock).  The published algorithm is restated from ``transformers`` 5.5.0
  ``VibeVoiceAcousticTokenizerDecoderModel`` and pinned against that module
  by the same fixture script.  Against the ONNX export itself: parity unpinned.

Reference line citations are relative to ``/root/reference/src/smalltts``;
``hf:`` means transformers/models/vibevoice_acoustic_tokenizer/
modeling_vibevoice_acoustic_tokenizer.py (transformers 5.5.0).
"""

from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

SAMPLE_RATE = 24_000  # infer/onnx.py:11
HOP_SIZE = 3_200  # infer/onnx.py:12
NUM_STEPS = 4  # infer/onnx.py:13
LATENT_DIM = 64
DIT_DIM = 960
DIT_HEADS = 8
DIT_HEAD_DIM = 120
DIT_BLOCKS = 12
ROPE_DIM = 64
VOC_RATIOS = (8, 5, 5, 4, 2, 2)  # hf config: reversed downsampling_ratios
VOC_DEPTHS = (8, 3, 3, 3, 3, 3, 3)  # hf config: reversed depths
VOC_FILTERS = 32


# --------------------------------------------------------------------------
# GEMM hook.  ``mode="fp32"`` is the oracle proper.  ``mode="bf16"`` rounds the
# two operands of every tensor-core contraction to bf16 (fp32 accumulate), which
# is what the CUDA engine's tcgen05 path does; tests use it to separate
# precision effects from kernel bugs.  It never changes the fp32 oracle.
# --------------------------------------------------------------------------
class _Mode:
    gemm = "fp32"


def set_gemm_mode(mode: str) -> None:
    assert mode in ("fp32", "bf16")
    _Mode.gemm = mode


def _r(x: Tensor) -> Tensor:
    if _Mode.gemm == "bf16":
        return x.to(torch.bfloat16).to(torch.float32)
    return x


def _linear(x: Tensor, w: Tensor, b: Optional[Tensor] = None) -> Tensor:
    y = _r(x) @ _r(w).t()
    if b is not None:
        y = y + b
    return y


# --------------------------------------------------------------------------
# schedule, RoPE tables, time embedding
# --------------------------------------------------------------------------
def alpha_sigma(t: float, eps: float = 1e-5):
    """infer/onnx.py:31-39 (== train/utils.py:12-22, server pipeline.rs:216-222)."""
    t = np.clip(t, eps, 1 - eps)
    a2 = np.cos(np.pi / 2 * t) ** 2
    log_snr = np.log(a2 / (1 - a2)) + 2 * np.log(0.5)
    alpha_sq = 1.0 / (1.0 + np.exp(-log_snr))
    return np.float32(np.sqrt(alpha_sq)), np.float32(np.sqrt(1 - alpha_sq))


def default_timesteps(num_steps: int = NUM_STEPS) -> np.ndarray:
    """infer/onnx.py:102: np.linspace(1, 0, NUM_STEPS, dtype=float32)."""
    return np.linspace(1, 0, num_steps, dtype=np.float32)


def rope_angles(seq_len: int, dim: int = ROPE_DIM) -> Tensor:
    """infer/onnx.py:42-47 == models/backbone/dit.py:138-149. (1, T, dim) angles,
    each frequency repeated for the two members of an interleaved pair."""
    inv_freq = 1.0 / (1e4 ** (torch.arange(0, dim, 2, dtype=torch.float32) / dim))
    t = torch.arange(seq_len, dtype=torch.float32)
    f = torch.einsum("i,j->ij", t, inv_freq)
    return torch.stack((f, f), dim=-1).reshape(1, seq_len, dim)


def _rotate_pairs(x: Tensor, ang: Tensor) -> Tensor:
    """dit.py:152-173 on the rotated slice: (x0,x1)->(x0 c - x1 s, x1 c + x0 s)."""
    x0, x1 = x[..., 0::2], x[..., 1::2]
    c, s = ang[..., 0::2].cos(), ang[..., 0::2].sin()
    return torch.stack((x0 * c - x1 * s, x1 * c + x0 * s), dim=-1).flatten(-2)


def time_embedding(sd: SD, t: Tensor) -> Tensor:
    """models/backbone/model.py:23-30."""
    half = 128
    f = torch.exp(torch.arange(half).float() * -(math.log(1e4) / (half - 1)))
    e = 1e3 * t.float()[:, None] * f[None, :]
    e = torch.cat((e.sin(), e.cos()), dim=-1)
    h = F.silu(_linear(e, sd["time_embedding.mlp.0.weight"], sd["time_embedding.mlp.0.bias"]))
    return _linear(h, sd["time_embedding.mlp.2.weight"], sd["time_embedding.mlp.2.bias"])


def rms_norm(x: Tensor, w: Tensor, eps: float) -> Tensor:
    """dit.py:42-53.  ``w`` is (d,) or (heads, d); normalise the last dim only."""
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * w


def _attend(q: Tensor, k: Tensor, v: Tensor, key_ok: Tensor) -> Tensor:
    """softmax(q k^T / sqrt(d) + keymask) v; q,k,v (B,H,N,d); key_ok (B,S) bool.
    Rows with no admissible key give 0 (the reference overwrites such rows)."""
    d = q.shape[-1]
    s = (_r(q) @ _r(k).transpose(-1, -2)) / math.sqrt(d)
    s = s.masked_fill(~key_ok[:, None, None, :], float("-inf"))
    m = s.amax(-1, keepdim=True)
    m = torch.where(torch.isfinite(m), m, torch.zeros_like(m))
    p = torch.exp(s - m)
    den = p.sum(-1, keepdim=True)
    p = p / torch.where(den > 0, den, torch.ones_like(den))
    return _r(p) @ _r(v)


# --------------------------------------------------------------------------
# encoders (style.py, phonemes.py)
# --------------------------------------------------------------------------
def _freqs_cis_angles(head_dim: int, n: int) -> Tensor:
    """style.py:13-18 / phonemes.py:72-79: angle[p, i] = p * 10000^(-2i/head_dim)."""
    inv = 1.0 / (10000.0 ** (torch.arange(0, head_dim, 2)[: head_dim // 2].float() / head_dim))
    return torch.outer(torch.arange(n).float(), inv)


def _encoder_block(sd: SD, pre: str, x: Tensor, key_ok: Tensor, ang: Tensor, heads: int, eps: float) -> Tensor:
    """style.py:28-105 == phonemes.py:87-167 (same block, different sizes)."""
    b, n, d = x.shape
    hd = d // heads
    h = rms_norm(x, sd[pre + "attention_norm.weight"], eps)
    q = _linear(h, sd[pre + "attention.wq.weight"]).reshape(b, n, heads, hd)
    k = _linear(h, sd[pre + "attention.wk.weight"]).reshape(b, n, heads, hd)
    v = _linear(h, sd[pre + "attention.wv.weight"]).reshape(b, n, heads, hd)
    gate = _linear(h, sd[pre + "attention.gate.weight"])
    q = rms_norm(q, sd[pre + "attention.q_norm.weight"], eps)
    k = rms_norm(k, sd[pre + "attention.k_norm.weight"], eps)
    # complex multiply by exp(i*angle) on interleaved pairs, full head dim
    a2 = ang[None, :, None, :].repeat_interleave(2, dim=-1)
    q = _rotate_pairs(q, a2)
    k = _rotate_pairs(k, a2)
    o = _attend(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), key_ok)
    o = o.transpose(1, 2).reshape(b, n, d) * torch.sigmoid(gate)
    x = x + _linear(o, sd[pre + "attention.wo.weight"])
    h = rms_norm(x, sd[pre + "mlp_norm.weight"], eps)
    m = F.silu(_linear(h, sd[pre + "mlp.w1.weight"])) * _linear(h, sd[pre + "mlp.w3.weight"])
    return x + _linear(m, sd[pre + "mlp.w2.weight"])


def style_encoder(sd: SD, ref: Tensor, ref_len: Tensor):
    """style.py:144-174: (B,R,64),(B,) -> (B,R,960), mask (B,R)."""
    b, r, _ = ref.shape
    mask = torch.arange(r)[None, :] < ref_len.clamp(max=r)[:, None]
    x = _linear(ref, sd["style_encoder.in_proj.weight"], sd["style_encoder.in_proj.bias"])
    x = x * sd["style_encoder.log_scale"].exp()
    ang = _freqs_cis_angles(64, r)
    for i in range(12):
        x = _encoder_block(sd, f"style_encoder.blocks.{i}.", x, mask, ang, 8, 1e-5)
    x = rms_norm(x, sd["style_encoder.norm.weight"], 1e-5)
    x = _linear(x, sd["style_encoder.out_proj.weight"], sd["style_encoder.out_proj.bias"])
    return x.masked_fill(~mask[..., None], 0.0), mask


def text_encoder(sd: SD, ids: Tensor, mask: Tensor) -> Tensor:
    """phonemes.py:200-207: (B,P) int64, (B,P) bool -> (B,P,512)."""
    x = sd["phoneme_embedding.text_embedding.weight"][ids]
    ang = _freqs_cis_angles(128, ids.shape[1])
    for i in range(8):
        x = _encoder_block(sd, f"phoneme_embedding.blocks.{i}.", x, mask, ang, 4, 1e-6)
    return rms_norm(x, sd["phoneme_embedding.norm.weight"], 1e-6)


def _heads(x: Tensor) -> Tensor:
    b, n, _ = x.shape
    return x.reshape(b, n, DIT_HEADS, DIT_HEAD_DIM)


def encode_conditions(sd: SD, ref: Tensor, ref_len: Tensor, phonemes: Tensor, phonemes_mask: Tensor) -> dict:
    """model.py:88-95 + dit.py:293-314 (== condition_encoder.onnx, infer/onnx.py:94).
    Returns per-block cross K/V (B,8,N,120) plus the two key masks."""
    ref_seq, ref_mask = style_encoder(sd, ref, ref_len)
    ph = text_encoder(sd, phonemes, phonemes_mask)
    mem = _linear(ph, sd["dit.phoneme_proj.weight"], sd["dit.phoneme_proj.bias"])
    mem = mem.masked_fill(~phonemes_mask[..., None], 0.0)
    layers = []
    for i in range(DIT_BLOCKS):
        p = f"dit.transformer_blocks.{i}.attn."
        wn = sd[p + "k_norm_cross.weight"]
        layer = {}
        for name, seq in (("ref", ref_seq), ("text", mem)):
            k = rms_norm(_heads(_linear(seq, sd[p + f"to_k_{name}.weight"], sd[p + f"to_k_{name}.bias"])), wn, 1e-6)
            v = _heads(_linear(seq, sd[p + f"to_v_{name}.weight"], sd[p + f"to_v_{name}.bias"]))
            layer["k_" + name] = k.transpose(1, 2)
            layer["v_" + name] = v.transpose(1, 2)
        layers.append(layer)
    return {"layers": layers, "ref_mask": ref_mask, "phonemes_mask": phonemes_mask}


# --------------------------------------------------------------------------
# denoiser (dit.py:209-253,316-327; model.py:97-100)
# --------------------------------------------------------------------------
def _layer_norm(x: Tensor) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), eps=1e-6)


def input_embedding(sd: SD, x: Tensor, mask: Tensor) -> Tensor:
    """dit.py:215-253: proj 64->960, then two grouped k=31 convs with Mish, + residual."""
    h = _linear(x, sd["dit.input_embed.proj.weight"], sd["dit.input_embed.proj.bias"])
    m3 = mask[..., None]
    c = h.masked_fill(~m3, 0.0).permute(0, 2, 1)
    p = "dit.input_embed.conv_pos_embed."
    c = F.mish(F.conv1d(_r(c), _r(sd[p + "conv1.weight"]), sd[p + "conv1.bias"], padding=15, groups=16)) * mask[:, None, :]
    c = F.mish(F.conv1d(_r(c), _r(sd[p + "conv2.weight"]), sd[p + "conv2.bias"], padding=15, groups=16))
    c = c.permute(0, 2, 1).masked_fill(~m3, 0.0)
    return c + h


def adaln_modulation(sd: SD, t: Tensor) -> dict:
    """Everything that depends on ``t`` only: model.py:23-30 -> dit.py:270-274,318 ->
    per-block dit.py:19-23 (chunk order shift_msa, scale_msa, gate_msa, shift_mlp,
    scale_mlp, gate_mlp) and final dit.py:36-37 (scale, shift)."""
    te = time_embedding(sd, t)
    e = _linear(F.silu(_linear(te, sd["dit.emb_proj.0.weight"], sd["dit.emb_proj.0.bias"])),
                sd["dit.emb_proj.2.weight"], sd["dit.emb_proj.2.bias"])
    se = F.silu(e)
    blocks = []
    for i in range(DIT_BLOCKS):
        p = f"dit.transformer_blocks.{i}.attn_norm.linear."
        blocks.append(_linear(se, sd[p + "weight"], sd[p + "bias"]).reshape(-1, 6, DIT_DIM))
    fin = _linear(se, sd["dit.norm_out.linear.weight"], sd["dit.norm_out.linear.bias"]).reshape(-1, 2, DIT_DIM)
    return {"blocks": blocks, "final": fin}


def dit_block(sd: SD, i: int, x: Tensor, mod: Tensor, mask: Tensor, layer: dict, key_ok: Tensor, ang: Tensor) -> Tensor:
    """dit.py:197-212 + JointAttention.forward_cached dit.py:95-119,131-135."""
    p = f"dit.transformer_blocks.{i}."
    shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = [mod[:, j][:, None, :] for j in range(6)]
    n = _layer_norm(x) * (1 + scale_msa) + shift_msa
    a = p + "attn."
    q = rms_norm(_heads(_linear(n, sd[a + "to_q.weight"], sd[a + "to_q.bias"])), sd[a + "q_norm.weight"], 1e-6)
    k = rms_norm(_heads(_linear(n, sd[a + "to_k_self.weight"], sd[a + "to_k_self.bias"])), sd[a + "k_norm.weight"], 1e-6)
    v = _heads(_linear(n, sd[a + "to_v_self.weight"], sd[a + "to_v_self.bias"]))
    gate = _linear(n, sd[a + "gate.weight"])
    a3 = ang[:, :, None, :]  # (1,T,1,64)
    q = torch.cat((_rotate_pairs(q[..., :ROPE_DIM], a3), q[..., ROPE_DIM:]), dim=-1).transpose(1, 2)
    k = torch.cat((_rotate_pairs(k[..., :ROPE_DIM], a3), k[..., ROPE_DIM:]), dim=-1).transpose(1, 2)
    v = v.transpose(1, 2)
    kk = torch.cat((k, layer["k_ref"], layer["k_text"]), dim=2)
    vv = torch.cat((v, layer["v_ref"], layer["v_text"]), dim=2)
    o = _attend(q, kk, vv, key_ok)
    o = o.transpose(1, 2).reshape(x.shape[0], x.shape[1], DIT_DIM) * torch.sigmoid(gate)
    o = _linear(o, sd[a + "to_out.0.weight"]).masked_fill(~mask[..., None], 0.0)
    x = x + torch.tanh(gate_msa) * o
    n2 = _layer_norm(x) * (1 + scale_mlp) + shift_mlp
    f = p + "ff."
    hdn = F.silu(_linear(n2, sd[f + "w1.weight"], sd[f + "w1.bias"])) * _linear(n2, sd[f + "w3.weight"], sd[f + "w3.bias"])
    ff = _linear(hdn, sd[f + "w2.weight"], sd[f + "w2.bias"])
    return x + torch.tanh(gate_mlp) * ff


def denoise_step(sd: SD, x_t: Tensor, mask: Tensor, t: Tensor, cond: dict,
                 taps: Optional[Callable[[str, Tensor], None]] = None) -> Tensor:
    """model.py:97-100 (== denoiser.onnx, infer/onnx.py:107-124): velocity (B,T,64)."""
    mod = adaln_modulation(sd, t)
    x = input_embedding(sd, x_t, mask)
    if taps:
        taps("input_embed", x)
    key_ok = torch.cat((mask, cond["ref_mask"], cond["phonemes_mask"]), dim=1)
    ang = rope_angles(x_t.shape[1])
    for i in range(DIT_BLOCKS):
        x = dit_block(sd, i, x, mod["blocks"][i], mask, cond["layers"][i], key_ok, ang)
        if taps:
            taps(f"block{i}", x)
    scale, shift = mod["final"][:, 0][:, None, :], mod["final"][:, 1][:, None, :]
    y = _layer_norm(x) * (1 + scale) + shift
    return _linear(y, sd["velocity.weight"], sd["velocity.bias"])


def sample(sd: SD, cond: dict, mask: Tensor, noise: Tensor, timesteps: Optional[Sequence[float]] = None) -> Tensor:
    """The DMD re-noising loop, infer/onnx.py:98-125, batched and with caller noise.
    noise: (steps,B,T,64).  Returns x_pred (B,T,64)."""
    ts = default_timesteps(noise.shape[0]) if timesteps is None else np.asarray(timesteps, dtype=np.float32)
    b = mask.shape[0]
    x_pred = torch.zeros_like(noise[0])
    for s, t_val in enumerate(ts):
        alpha, sigma = alpha_sigma(float(t_val))
        x_t = float(alpha) * x_pred + float(sigma) * noise[s]
        v = denoise_step(sd, x_t, mask, torch.full((b,), float(t_val)), cond)
        x_pred = float(alpha) * x_t - float(sigma) * v
    return x_pred


def cfg_conditions(sd: SD, ref: Tensor, ref_len: Tensor, phonemes: Tensor, phonemes_mask: Tensor) -> dict:
    """Conditions of the teacher's 3-way CFG batch, rows [cond | text dropped | speaker dropped]
    (scripts/train/dmd2/distill.py:74-96: dropped text = zero ids + all-False mask, dropped speaker = zero latents
    + zero length)."""
    ref3 = torch.cat([ref, ref, torch.zeros_like(ref)], dim=0)
    len3 = torch.cat([ref_len, ref_len, torch.zeros_like(ref_len)], dim=0)
    ph3 = torch.cat([phonemes, torch.zeros_like(phonemes), phonemes], dim=0)
    pm3 = torch.cat([phonemes_mask, torch.zeros_like(phonemes_mask), phonemes_mask], dim=0)
    return encode_conditions(sd, ref3, len3, ph3, pm3)


def cfg_velocity(sd: SD, x_t: Tensor, mask: Tensor, t: Tensor, cond3: dict, cfg_scale_text: float = 2.0,
                 cfg_scale_speaker: float = 1.5) -> Tensor:
    """distill.py:75,92-103: one 3B-row evaluation, velocity = v_c + s_t (v_c - v_ut) + s_s (v_c - v_us).
    (model(...) there is the uncached forward, which equals encode_conditions + denoise_step bit for bit.)"""
    v3 = denoise_step(sd, x_t.repeat(3, 1, 1), mask.repeat(3, 1), t.repeat(3), cond3)
    v_c, v_ut, v_us = v3.chunk(3, dim=0)
    return v_c + cfg_scale_text * (v_c - v_ut) + cfg_scale_speaker * (v_c - v_us)


def sample_teacher(sd: SD, cond3: dict, mask: Tensor, noise: Tensor, steps: int = 128, cfg_scale_text: float = 2.0,
                   cfg_scale_speaker: float = 1.5) -> Tensor:
    """Teacher sampler for BASELINE config 5.  The reference has no teacher inference script; this is the sampler
    its training code implies: t = linspace(1, 0, steps + 1), CFG velocity as above, and the
    v-prediction identities of train/utils.py:54-67 / distill.py:127-130:
        x0 = alpha x_t - sigma v,  eps = sigma x_t + alpha v,  x_{t'} = alpha' x0 + sigma' eps.
    noise: (B,T,64) = x_1.  Returns x at t = 0."""
    ts = np.linspace(1.0, 0.0, steps + 1).astype(np.float32)
    b = mask.shape[0]
    x = noise.clone()
    for s in range(steps):
        a, sg = alpha_sigma(float(ts[s]))
        an, sn = alpha_sigma(float(ts[s + 1]))
        v = cfg_velocity(sd, x, mask, torch.full((b,), float(ts[s])), cond3, cfg_scale_text, cfg_scale_speaker)
        x0 = float(a) * x - float(sg) * v
        eps = float(sg) * x + float(a) * v
        x = float(an) * x0 + float(sn) * eps
    return x


# --------------------------------------------------------------------------
# vocoder = VibeVoice acoustic-tokenizer decoder (hf:181-297,406-500)
# --------------------------------------------------------------------------
def _causal_conv(x: Tensor, w: Tensor, b: Tensor, groups: int = 1) -> Tensor:
    """hf:181-216, stride 1 / dilation 1: left-pad k-1 zeros."""
    k = w.shape[-1]
    wr = w if groups > 1 else _r(w)
    xr = x if groups > 1 else _r(x)
    return F.conv1d(F.pad(xr, (k - 1, 0)), wr, b, groups=groups)


def _causal_convtr(x: Tensor, w: Tensor, b: Tensor, stride: int) -> Tensor:
    """hf:219-260: ConvTranspose1d(k=2r, stride=r) then drop the last k-r samples."""
    y = F.conv_transpose1d(_r(x), _r(w), b, stride=stride)
    return y[..., : -(w.shape[-1] - stride)]


def _convnext(sd: SD, pre: str, x: Tensor) -> Tensor:
    """hf:263-297.  x is (B,C,T)."""
    c = x.shape[1]
    h = rms_norm(x.transpose(1, 2), sd[pre + "norm.weight"], 1e-5).transpose(1, 2)
    h = _causal_conv(h, sd[pre + "mixer.conv.weight"], sd[pre + "mixer.conv.bias"], groups=c)
    x = x + h * sd[pre + "gamma"][:, None]
    h = rms_norm(x.transpose(1, 2), sd[pre + "ffn_norm.weight"], 1e-5)
    h = F.gelu(_linear(h, sd[pre + "ffn.linear1.weight"], sd[pre + "ffn.linear1.bias"]))
    h = _linear(h, sd[pre + "ffn.linear2.weight"], sd[pre + "ffn.linear2.bias"]).transpose(1, 2)
    return x + h * sd[pre + "ffn_gamma"][:, None]


def vocoder_decode(vsd: SD, latents: Tensor, taps: Optional[Callable[[str, Tensor], None]] = None) -> Tensor:
    """codec/onnx.py:42-53 (decoder.onnx) per hf:406-500,538-548: (B,T,64)->(B,1,T*3200)."""
    x = latents.permute(0, 2, 1)
    x = _causal_conv(x, vsd["stem.conv.conv.weight"], vsd["stem.conv.conv.bias"])
    for l in range(VOC_DEPTHS[0]):
        x = _convnext(vsd, f"stem.stage.{l}.", x)
    if taps:
        taps("stem", x)
    for s, r in enumerate(VOC_RATIOS):
        x = _causal_convtr(x, vsd[f"conv_layers.{s}.convtr.convtr.weight"], vsd[f"conv_layers.{s}.convtr.convtr.bias"], r)
        for l in range(VOC_DEPTHS[s + 1]):
            x = _convnext(vsd, f"conv_layers.{s}.stage.{l}.", x)
        if taps:
            taps(f"up{s}", x)
    return _causal_conv(x, vsd["head.conv.weight"], vsd["head.conv.bias"])


# --------------------------------------------------------------------------
# whole path (infer/onnx.py:68-129), batched with ragged lengths
# --------------------------------------------------------------------------
ENC_RATIOS = (2, 2, 4, 5, 5, 8)  # hf encoder config
ENC_DEPTHS = (3, 3, 3, 3, 3, 3, 8)


def _causal_conv_strided(x: Tensor, w: Tensor, b: Tensor, stride: int) -> Tensor:
    """hf:181-216 with stride r and k = 2r: left-pad (k-1) - (r-1) = r zeros, then Conv1d(stride=r); T -> T // r."""
    k = w.shape[-1]
    return F.conv1d(F.pad(_r(x), (k - 1 - (stride - 1), 0)), _r(w), b, stride=stride)


def codec_encode(esd: SD, audio: Tensor) -> Tensor:
    """codec/onnx.py:56-75 == encoder.onnx; arithmetic per hf:300-403 (VibeVoiceAcousticTokenizerEncoderModel):
    stem CausalConv1d(1->32,k7) + 3 ConvNeXt layers; six [strided CausalConv1d(C->2C, k=2r, stride r) + ConvNeXt
    layers] with r = 2,2,4,5,5,8; head CausalConv1d(2048->64,k7).  audio (B,1,N) -> latents (B, N // 3200, 64)
    (the VAE mean; whether the ONNX export adds sampling noise is unknown)."""
    x = _causal_conv(audio, esd["stem.conv.conv.weight"], esd["stem.conv.conv.bias"])
    for l in range(ENC_DEPTHS[0]):
        x = _convnext(esd, f"stem.stage.{l}.", x)
    for i, r in enumerate(ENC_RATIOS):
        x = _causal_conv_strided(x, esd[f"conv_layers.{i}.conv.conv.weight"], esd[f"conv_layers.{i}.conv.conv.bias"], r)
        for l in range(ENC_DEPTHS[i + 1]):
            x = _convnext(esd, f"conv_layers.{i}.stage.{l}.", x)
    x = _causal_conv(x, esd["head.conv.weight"], esd["head.conv.bias"])
    return x.permute(0, 2, 1)


# ---------------------------------------------------------------------------------------------- resample_hq
def resample_bank(sr_from: int, sr_to: int, lowpass_filter_width: int = 1024, rolloff: float = 0.94,
                  beta: float = 14.769656459379492):
    """Polyphase filter bank of the reference's ``make_resampler`` (infer/utils.py:8-16): torchaudio's
    ``_get_sinc_resample_kernel`` for ``sinc_interp_kaiser``, restated.  Returns (bank fp32 [up, K], width, down, up).

    bank[j, k] = sinc(pi t) * kaiser(t) * base/down with t = clamp((-j/up + (k - width)/down) * base, +-lpw),
    base = min(down, up) * rolloff, width = ceil(lpw * down / base), K = 2 width + down; fp64 arithmetic rounded to
    fp32 at the end, except that torchaudio forms -j/up from an integer tensor (fp32 division) and holds beta in a
    fp32 tensor."""
    g = math.gcd(int(sr_from), int(sr_to))
    down, up = int(sr_from) // g, int(sr_to) // g
    base = min(down, up) * rolloff
    width = math.ceil(lowpass_filter_width * down / base)
    k = np.arange(-width, width + down, dtype=np.float64) / down
    phase = (np.arange(0, -up, -1, dtype=np.float32) / np.float32(up)).astype(np.float64)
    t = (phase[:, None] + k[None, :]) * base
    t = np.clip(t, -lowpass_filter_width, lowpass_filter_width)
    b32 = torch.tensor(float(beta))  # fp32, like torchaudio's beta_tensor
    win = torch.i0(b32.double() * torch.sqrt(1 - torch.from_numpy(t / lowpass_filter_width) ** 2)) / torch.i0(b32).double()
    t = t * math.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        sinc = np.where(t == 0, 1.0, np.sin(t) / t)
    bank = sinc * win.numpy() * (base / down)
    return bank.astype(np.float32), width, down, up


def resample_hq(x: np.ndarray, sr: int, target: int) -> np.ndarray:
    """infer/utils.py:19-23 (torchaudio ``_apply_sinc_resample_kernel``): x fp32 [B, N] at ``sr`` -> [B, ceil(N *
    target / sr)] at ``target``.  out[n*up + j] = sum_k bank[j, k] * xpad[n*down + k], xpad = x with ``width`` zeros in
    front and ``width + down`` behind; products accumulated in fp64, result rounded to fp32."""
    x = np.asarray(x, dtype=np.float32)
    if sr == target:
        return x
    bank, width, down, up = resample_bank(sr, target)
    B, N = x.shape
    K = bank.shape[1]
    xp = np.pad(x.astype(np.float64), ((0, 0), (width, width + down)))
    win = np.lib.stride_tricks.sliding_window_view(xp, K, axis=1)[:, ::down]  # [B, frames, K]
    y = np.einsum("bfk,jk->bfj", win, bank.astype(np.float64)).reshape(B, -1)
    n_out = -((-up * N) // down)
    return y[:, :n_out].astype(np.float32)


def frames_for(duration_sec: float) -> int:
    """infer/onnx.py:84."""
    return max(1, int(duration_sec * SAMPLE_RATE / HOP_SIZE))


def pad_batch(ref_list: Sequence[Tensor], ids_list: Sequence[Sequence[int]], frames: Sequence[int]):
    b = len(ref_list)
    r_max = max(int(r.shape[0]) for r in ref_list)
    p_max = max(len(p) for p in ids_list)
    t_max = max(frames)
    ref = torch.zeros(b, r_max, LATENT_DIM)
    ref_len = torch.zeros(b, dtype=torch.int64)
    ids = torch.zeros(b, p_max, dtype=torch.int64)
    pmask = torch.zeros(b, p_max, dtype=torch.bool)
    mask = torch.zeros(b, t_max, dtype=torch.bool)
    for i in range(b):
        n = int(ref_list[i].shape[0])
        ref[i, :n] = torch.as_tensor(ref_list[i], dtype=torch.float32)
        ref_len[i] = n
        ids[i, : len(ids_list[i])] = torch.as_tensor(list(ids_list[i]), dtype=torch.int64)
        pmask[i, : len(ids_list[i])] = True
        mask[i, : frames[i]] = True
    return ref, ref_len, ids, pmask, mask


@torch.inference_mode()
def synthesize_batch(sd: SD, vsd: SD, ref_list, ids_list, frames: Sequence[int], noise: Tensor,
                     timesteps: Optional[Sequence[float]] = None) -> List[Tensor]:
    """Batched SmallTTS.synthesize.  noise: (steps,B,Tmax,64).  Returns per-utterance
    (1, frames_i*3200) audio; row i equals a batch-1 run of utterance i (masks
    isolate rows, causal vocoder ignores right padding)."""
    ref, ref_len, ids, pmask, mask = pad_batch(ref_list, ids_list, frames)
    cond = encode_conditions(sd, ref, ref_len, ids, pmask)
    lat = sample(sd, cond, mask, noise, timesteps)
    audio = vocoder_decode(vsd, lat)
    return [audio[i, :, : frames[i] * HOP_SIZE] for i in range(len(frames))]

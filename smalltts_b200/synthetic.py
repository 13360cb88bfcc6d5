"""Seeded random weights with the reference's state-dict key names and shapes.

There is no network in the build or benchmark environment, so neither the
smalltts DMD checkpoint nor the VibeVoice codec weights can be fetched.  The
engine, the parity tests and ``bench.py`` therefore run on weights drawn here:
same architecture, same tensor names (``DiTModel.state_dict()`` of
models/backbone/model.py:33-54 and the HF ``VibeVoiceAcousticTokenizerDecoderModel``),
deterministic for a given seed on a given torch build.

The reference zero-initialises every adaLN projection and the velocity head
(models/backbone/dit.py:281-285, model.py:53-54) and VibeVoice starts its layer
scales at 1e-6; with those values the network output is identically ~0 and a
parity test would be vacuous, so these tensors are drawn non-trivially.
"""

from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch

Spec = Tuple[str, Tuple[int, ...], str, float]


def _enc_block(prefix: str, d: int, heads: int, inter: int) -> List[Spec]:
    hd = d // heads
    out: List[Spec] = []
    for w in ("wq", "wk", "wv", "wo", "gate"):
        out.append((f"{prefix}attention.{w}.weight", (d, d), "lin", d))
    out.append((f"{prefix}attention.q_norm.weight", (heads, hd), "norm", 0))
    out.append((f"{prefix}attention.k_norm.weight", (heads, hd), "norm", 0))
    out.append((f"{prefix}mlp.w1.weight", (inter, d), "lin", d))
    out.append((f"{prefix}mlp.w3.weight", (inter, d), "lin", d))
    out.append((f"{prefix}mlp.w2.weight", (d, inter), "lin", inter))
    out.append((f"{prefix}attention_norm.weight", (d,), "norm", 0))
    out.append((f"{prefix}mlp_norm.weight", (d,), "norm", 0))
    return out


def dit_specs() -> List[Spec]:
    """SURVEY.md appendix E / DiTModel(64).state_dict(): 592 tensors, 327,756,609 params."""
    D, H, HD, FF = 960, 8, 120, 2400
    s: List[Spec] = [
        ("time_embedding.mlp.0.weight", (D, 256), "lin", 256),
        ("time_embedding.mlp.0.bias", (D,), "lin", 256),
        ("time_embedding.mlp.2.weight", (D, D), "lin", D),
        ("time_embedding.mlp.2.bias", (D,), "lin", D),
        ("phoneme_embedding.text_embedding.weight", (198, 512), "emb", 0),
    ]
    for i in range(8):
        s += _enc_block(f"phoneme_embedding.blocks.{i}.", 512, 4, 1024)
    s.append(("phoneme_embedding.norm.weight", (512,), "norm", 0))
    s += [
        ("style_encoder.log_scale", (), "const", -1.8),
        ("style_encoder.in_proj.weight", (512, 64), "lin", 64),
        ("style_encoder.in_proj.bias", (512,), "lin", 64),
    ]
    for i in range(12):
        s += _enc_block(f"style_encoder.blocks.{i}.", 512, 8, 1536)
    s += [
        ("style_encoder.norm.weight", (512,), "norm", 0),
        ("style_encoder.out_proj.weight", (D, 512), "lin", 512),
        ("style_encoder.out_proj.bias", (D,), "lin", 512),
        ("dit.input_embed.proj.weight", (D, 64), "lin", 64),
        ("dit.input_embed.proj.bias", (D,), "lin", 64),
    ]
    for c in ("conv1", "conv2"):
        s.append((f"dit.input_embed.conv_pos_embed.{c}.weight", (D, 60, 31), "lin", 60 * 31))
        s.append((f"dit.input_embed.conv_pos_embed.{c}.bias", (D,), "lin", 60 * 31))
    s += [
        ("dit.phoneme_proj.weight", (D, 512), "lin", 512),
        ("dit.phoneme_proj.bias", (D,), "lin", 512),
        ("dit.emb_proj.0.weight", (2 * D, D), "lin", D),
        ("dit.emb_proj.0.bias", (2 * D,), "lin", D),
        ("dit.emb_proj.2.weight", (D, 2 * D), "lin", 2 * D),
        ("dit.emb_proj.2.bias", (D,), "lin", 2 * D),
    ]
    for i in range(12):
        p = f"dit.transformer_blocks.{i}."
        s.append((p + "attn_norm.linear.weight", (6 * D, D), "normal", 0.02))
        s.append((p + "attn_norm.linear.bias", (6 * D,), "normal", 0.02))
        for w in ("to_q", "to_k_self", "to_v_self", "to_k_ref", "to_v_ref", "to_k_text", "to_v_text"):
            s.append((f"{p}attn.{w}.weight", (D, D), "lin", D))
            s.append((f"{p}attn.{w}.bias", (D,), "lin", D))
        s.append((p + "attn.gate.weight", (D, D), "lin", D))
        s.append((p + "attn.to_out.0.weight", (D, D), "lin", D))
        for w in ("q_norm", "k_norm", "k_norm_cross"):
            s.append((f"{p}attn.{w}.weight", (H, HD), "norm", 0))
        for w in ("w1", "w3"):
            s.append((f"{p}ff.{w}.weight", (FF, D), "lin", D))
            s.append((f"{p}ff.{w}.bias", (FF,), "lin", D))
        s.append((p + "ff.w2.weight", (D, FF), "lin", FF))
        s.append((p + "ff.w2.bias", (D,), "lin", FF))
    s += [
        ("dit.norm_out.linear.weight", (2 * D, D), "normal", 0.02),
        ("dit.norm_out.linear.bias", (2 * D,), "normal", 0.02),
        ("velocity.weight", (64, D), "normal", 0.02),
        ("velocity.bias", (64,), "normal", 0.02),
    ]
    return s


VOC_RATIOS = (8, 5, 5, 4, 2, 2)
VOC_DEPTHS = (8, 3, 3, 3, 3, 3, 3)
VOC_CHANNELS = (2048, 1024, 512, 256, 128, 64, 32)


def _convnext_specs(prefix: str, c: int) -> List[Spec]:
    return [
        (prefix + "gamma", (c,), "gamma", 0),
        (prefix + "ffn_gamma", (c,), "gamma", 0),
        (prefix + "norm.weight", (c,), "norm", 0),
        (prefix + "ffn_norm.weight", (c,), "norm", 0),
        (prefix + "ffn.linear1.weight", (4 * c, c), "lin", c),
        (prefix + "ffn.linear1.bias", (4 * c,), "lin", c),
        (prefix + "ffn.linear2.weight", (c, 4 * c), "lin", 4 * c),
        (prefix + "ffn.linear2.bias", (c,), "lin", 4 * c),
        (prefix + "mixer.conv.weight", (c, 1, 7), "lin", 7),
        (prefix + "mixer.conv.bias", (c,), "lin", 7),
    ]


def vocoder_specs() -> List[Spec]:
    """HF VibeVoiceAcousticTokenizerDecoderModel.state_dict(): 276 tensors, 343,695,969 params."""
    s: List[Spec] = [
        ("stem.conv.conv.weight", (2048, 64, 7), "lin", 64 * 7),
        ("stem.conv.conv.bias", (2048,), "lin", 64 * 7),
    ]
    for l in range(VOC_DEPTHS[0]):
        s += _convnext_specs(f"stem.stage.{l}.", 2048)
    for i, r in enumerate(VOC_RATIOS):
        cin, cout = VOC_CHANNELS[i], VOC_CHANNELS[i + 1]
        s.append((f"conv_layers.{i}.convtr.convtr.weight", (cin, cout, 2 * r), "lin", 2 * cin))
        s.append((f"conv_layers.{i}.convtr.convtr.bias", (cout,), "lin", 2 * cin))
        for l in range(VOC_DEPTHS[i + 1]):
            s += _convnext_specs(f"conv_layers.{i}.stage.{l}.", cout)
    s += [("head.conv.weight", (1, 32, 7), "lin", 32 * 7), ("head.conv.bias", (1,), "lin", 32 * 7)]
    return s


ENC_RATIOS = (2, 2, 4, 5, 5, 8)  # hf encoder config: downsampling_ratios
ENC_DEPTHS = (3, 3, 3, 3, 3, 3, 8)
ENC_CHANNELS = (32, 64, 128, 256, 512, 1024, 2048)


def encoder_specs() -> List[Spec]:
    """HF VibeVoiceAcousticTokenizerEncoderModel.state_dict() (hf:300-403): mirror of the decoder with strided
    causal convolutions."""
    s: List[Spec] = [
        ("stem.conv.conv.weight", (32, 1, 7), "lin", 7),
        ("stem.conv.conv.bias", (32,), "lin", 7),
    ]
    for l in range(ENC_DEPTHS[0]):
        s += _convnext_specs(f"stem.stage.{l}.", 32)
    for i, r in enumerate(ENC_RATIOS):
        cin, cout = ENC_CHANNELS[i], ENC_CHANNELS[i + 1]
        s.append((f"conv_layers.{i}.conv.conv.weight", (cout, cin, 2 * r), "lin", 2 * r * cin))
        s.append((f"conv_layers.{i}.conv.conv.bias", (cout,), "lin", 2 * r * cin))
        for l in range(ENC_DEPTHS[i + 1]):
            s += _convnext_specs(f"conv_layers.{i}.stage.{l}.", cout)
    s += [("head.conv.weight", (64, 2048, 7), "lin", 2048 * 7), ("head.conv.bias", (64,), "lin", 2048 * 7)]
    return s


def _draw(specs: List[Spec], seed: int) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for name, shape, kind, arg in specs:
        if kind == "lin":  # nn.Linear / nn.Conv default scale: U(+-1/sqrt(fan_in))
            t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(arg)
        elif kind == "norm":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind == "emb":
            t = torch.randn(shape, generator=g)
        elif kind == "normal":
            t = arg * torch.randn(shape, generator=g)
        elif kind == "gamma":
            t = 0.1 * (1.0 + torch.rand(shape, generator=g))
        elif kind == "const":
            t = torch.tensor(arg)
        else:  # pragma: no cover
            raise ValueError(kind)
        sd[name] = t.float().contiguous()
    return sd


def dit_state_dict(seed: int = 0) -> Dict[str, torch.Tensor]:
    return _draw(dit_specs(), seed)


def vocoder_state_dict(seed: int = 1) -> Dict[str, torch.Tensor]:
    return _draw(vocoder_specs(), seed)


def encoder_state_dict(seed: int = 2) -> Dict[str, torch.Tensor]:
    return _draw(encoder_specs(), seed)


def synthetic_inputs(batch: int, frames, ref_frames, n_phonemes, seed: int = 20260217, steps: int = 4):
    """SURVEY.md 8(d) synthetic inputs: ref latents ~ N(0,1), ids ~ U{1..197}, noise ~ N(0,1)
    from ``torch.Generator(seed)`` (seed = first seed in readme_samples/.../manifest.json)."""
    g = torch.Generator().manual_seed(seed)

    def per(v):
        return [int(v)] * batch if isinstance(v, int) else [int(x) for x in v]

    frames, ref_frames, n_phonemes = per(frames), per(ref_frames), per(n_phonemes)
    refs = [torch.randn(r, 64, generator=g) for r in ref_frames]
    ids = [torch.randint(1, 198, (p,), generator=g).tolist() for p in n_phonemes]
    noise = torch.randn(steps, batch, max(frames), 64, generator=g)
    return refs, ids, frames, noise

"""Generate the ONNX fixtures for the weight loader (test infrastructure; authoring container only).

    python oracle/make_golden_onnx.py            # tests/golden/tiny_codec_decoder.onnx + .npz
    python oracle/make_golden_onnx.py --full     # additionally: export the reference's full denoiser to /tmp and
                                                 # check that smalltts_b200.weights recovers every tensor (not committed)

The reference loads its weights from ONNX graphs (infer/onnx.py:60-63, codec/onnx.py:28-31) that are not in the
repository; what IS known is how such graphs are produced: ``torch.onnx.export`` of the PyTorch modules.  This
script runs torch's TorchScript ONNX exporter (no ``onnx`` package needed once its onnxscript post-pass is stubbed
out) on

* a scaled-down HF ``VibeVoiceAcousticTokenizerDecoderModel`` (same module tree / tensor names as the real
  decoder, 2 stages) -> committed fixture, a few hundred KB;
* with ``--full``: the reference's own ``DiTModel.denoise_step`` (models/backbone/model.py:97-100) on the seeded
  weights -> 1 GB in /tmp, validated and deleted.

The exporter folds nn.Linear weights into anonymous transposed ``onnx::MatMul_*`` initialisers, which is the case
the loader's graph walk exists for.
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

from torch.onnx._internal.torchscript_exporter import onnx_proto_utils  # noqa: E402

onnx_proto_utils._add_onnxscript_fn = lambda model_bytes, custom_opsets: model_bytes  # needs `onnx`; no-op for us

OUT = os.path.join(ROOT, "tests", "golden")


def export(module, args, path, inputs, outputs, dyn):
    torch.onnx.export(module, args, path, dynamo=False, opset_version=17, input_names=inputs, output_names=outputs,
                      dynamic_axes=dyn, do_constant_folding=True)


def tiny_decoder():
    from transformers.models.vibevoice_acoustic_tokenizer.configuration_vibevoice_acoustic_tokenizer import (
        VibeVoiceAcousticTokenizerConfig,
    )
    from transformers.models.vibevoice_acoustic_tokenizer.modeling_vibevoice_acoustic_tokenizer import (
        VibeVoiceAcousticTokenizerDecoderModel,
    )

    cfg = VibeVoiceAcousticTokenizerConfig(hidden_size=8, num_filters=4, downsampling_ratios=[2, 2], depths=[1, 1, 2])
    torch.manual_seed(7)
    m = VibeVoiceAcousticTokenizerDecoderModel(cfg.decoder_config).eval()
    for p in m.parameters():  # layer scales start at 1e-6: make every tensor distinguishable
        p.data = torch.randn_like(p) * 0.1
    x = torch.randn(1, 3, 8)

    class Wrap(torch.nn.Module):  # the reference feeds (B,T,C) and gets (B,1,N) (codec/onnx.py:42-53)
        def __init__(self, d):
            super().__init__()
            self.decoder = d

        def forward(self, latents):
            return self.decoder(latents.transpose(1, 2)).audio

    path = os.path.join(OUT, "tiny_codec_decoder.onnx")
    export(Wrap(m), (x,), path, ["latents"], ["audio"], {"latents": {0: "b", 1: "t"}})
    sd = {k: v.numpy() for k, v in m.state_dict().items()}
    np.savez(os.path.join(OUT, "tiny_codec_decoder.npz"), **sd)
    print("wrote", path, os.path.getsize(path), "bytes,", len(sd), "tensors")


def full_denoiser():
    from smalltts.models.backbone.model import DiTModel
    from smalltts_b200 import synthetic, weights

    import torch.nn.functional as F

    def rms_norm(x, shape, weight=None, eps=None):  # aten::rms_norm has no symbolic in the TorchScript exporter
        eps = torch.finfo(x.dtype).eps if eps is None else eps
        y = x * torch.rsqrt(x.pow(2).mean(tuple(range(-len(shape), 0)), keepdim=True) + eps)
        return y if weight is None else y * weight

    F.rms_norm = rms_norm
    model = DiTModel(64)
    sd = synthetic.dit_state_dict(0)
    model.load_state_dict(sd, strict=True)
    model.eval()
    B, T, R, P = 1, 6, 5, 7
    ref = torch.randn(B, R, 64)
    ids = torch.randint(1, 198, (B, P))
    cached = model.encode_conditions(ref, torch.tensor([R]), ids, torch.ones(B, P, dtype=torch.bool), T)

    class Den(torch.nn.Module):
        def __init__(self, m, cached):
            super().__init__()
            self.m, self.c = m, cached

        def forward(self, x_t, mask, t):
            return self.m.denoise_step(x_t, mask, t, self.c)

    path = "/tmp/stts_denoiser_full.onnx"
    export(Den(model, cached), (torch.randn(B, T, 64), torch.ones(B, T, dtype=torch.bool), torch.tensor([0.5])), path,
           ["x_t", "mask", "t"], ["velocity"], {})
    print("exported", os.path.getsize(path) / 1e6, "MB")
    got = weights.onnx_state_dict(path, synthetic.dit_specs(), weights.dit_exec_rank)
    anon = [k for k in weights.read_onnx(path).tensors if k.startswith("onnx::")]
    print(len(got), "tensors recovered;", len(anon), "anonymous initialisers in the file")
    # the denoiser graph holds the tensors of time embedding + dit (not the two condition encoders)
    want = [s for s in synthetic.dit_specs() if s[0].startswith(("time_embedding.", "dit.", "velocity."))
            and ".to_k_ref." not in s[0] and ".to_v_ref." not in s[0] and ".to_k_text." not in s[0]
            and ".to_v_text." not in s[0] and "k_norm_cross" not in s[0] and "phoneme_proj" not in s[0]]
    sel = weights.match_to_specs(got, want, "denoiser.onnx")
    worst = max(float(np.abs(np.asarray(sel[n]) - sd[n].numpy()).max()) for n, *_ in want)
    print(f"{len(want)} denoiser tensors matched by name and shape, max |diff| = {worst:.3g}")
    os.remove(path)


if __name__ == "__main__":
    torch.set_grad_enabled(False)
    tiny_decoder()
    if "--full" in sys.argv:
        full_denoiser()

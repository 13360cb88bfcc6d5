"""Weight readers (smalltts_b200/weights.py): ONNX initialisers without the onnx package, .pt / .safetensors /
.sttsw state dicts, and the name/shape check against the architecture.  CPU only."""
import os
import struct

import numpy as np
import pytest
import torch

from smalltts_b200 import synthetic, weights

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


# ---------------------------------------------------------------- a minimal protobuf writer (test side only)
def _vi(x: int) -> bytes:
    x &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = x & 0x7F
        x >>= 7
        out.append(b | (0x80 if x else 0))
        if not x:
            return bytes(out)


def _ld(fno: int, payload: bytes) -> bytes:
    return _vi(fno << 3 | 2) + _vi(len(payload)) + payload


def _v(fno: int, x: int) -> bytes:
    return _vi(fno << 3) + _vi(x)


def _tensor(name, dims, dtype, raw=None, float_data=None, int_data=None, external=None) -> bytes:
    b = b"".join(_v(1, d) for d in dims) + _v(2, dtype) + _ld(8, name.encode())
    if raw is not None:
        b += _ld(9, raw)
    if float_data is not None:
        b += _ld(4, np.asarray(float_data, "<f4").tobytes())
    if int_data is not None:
        b += _ld(5, b"".join(_vi(int(x)) for x in int_data))
    for k, val in (external or {}).items():
        b += _ld(13, _ld(1, k.encode()) + _ld(2, str(val).encode()))
    if external:
        b += _v(14, 1)
    return b


def _node(name, op, ins, outs, value_tensor=None) -> bytes:
    b = b"".join(_ld(1, i.encode()) for i in ins) + b"".join(_ld(2, o.encode()) for o in outs)
    b += _ld(3, name.encode()) + _ld(4, op.encode())
    if value_tensor is not None:
        b += _ld(5, _ld(1, b"value") + _ld(5, value_tensor) + _v(20, 4))
    return b


def _model(nodes, inits) -> bytes:
    graph = b"".join(_ld(1, n) for n in nodes) + _ld(2, b"g") + b"".join(_ld(5, t) for t in inits)
    return _v(1, 8) + _ld(2, b"pytest") + _ld(7, graph)


def test_onnx_reader_handcrafted(tmp_path):
    rng = np.random.default_rng(0)
    w = rng.standard_normal((3, 5)).astype(np.float32)  # Linear(5 -> 3), stored transposed by the exporter
    bias = rng.standard_normal(3).astype(np.float32)
    emb = rng.standard_normal((7, 4)).astype(np.float32)
    ext = rng.standard_normal((2, 6)).astype(np.float32)
    half = rng.standard_normal(4).astype(np.float16)
    bf = np.array([1.0, -2.5, 0.15625], dtype=np.float32)
    (tmp_path / "blob.bin").write_bytes(b"\0" * 16 + ext.tobytes())
    inits = [
        _tensor("onnx::MatMul_12", [5, 3], 1, raw=np.ascontiguousarray(w.T).tobytes()),
        _tensor("net.proj.bias", [3], 1, float_data=bias),
        _tensor("onnx::Gather_3", [7, 4], 1, raw=emb.tobytes()),
        _tensor("net.ext.weight", [2, 6], 1, external={"location": "blob.bin", "offset": 16, "length": ext.nbytes}),
        _tensor("net.h", [4], 10, int_data=half.view(np.uint16)),
        _tensor("net.b", [3], 16, raw=(bf.view(np.uint32) >> 16).astype("<u2").tobytes()),
        _tensor("net.count", [2], 7, raw=np.array([3, -1], "<i8").tobytes()),
    ]
    const = _tensor("", [2, 2], 1, raw=np.eye(2, dtype=np.float32).tobytes())
    nodes = [
        _node("/net/emb/Gather", "Gather", ["onnx::Gather_3", "ids"], ["e"]),
        _node("/net/proj/MatMul", "MatMul", ["e", "onnx::MatMul_12"], ["y"]),
        _node("/net/Constant", "Constant", [], ["/net/Constant_output_0"], value_tensor=const),
    ]
    path = tmp_path / "m.onnx"
    path.write_bytes(_model(nodes, inits))
    g = weights.read_onnx(str(path))
    assert [n[1] for n in g.nodes] == ["Gather", "MatMul", "Constant"]
    sd = weights.onnx_state_dict(str(path))
    np.testing.assert_array_equal(sd["net.proj.weight"], w)
    np.testing.assert_array_equal(sd["net.proj.bias"], bias)
    np.testing.assert_array_equal(sd["net.emb.weight"], emb)
    np.testing.assert_array_equal(sd["net.ext.weight"], ext)
    np.testing.assert_array_equal(sd["net.h"], half)
    np.testing.assert_array_equal(sd["net.b"], bf)
    np.testing.assert_array_equal(sd["net.count"], [3, -1])
    np.testing.assert_array_equal(g.tensors["/net/Constant_output_0"], np.eye(2))


def test_onnx_scopeless_graph_bias_pairing_and_execution_order(tmp_path):
    """A graph traced through direct method calls has node names like '/to_q/MatMul', '/gate_1/MatMul' (seen on a
    TorchScript export of the reference's DiTModel.denoise_step): Linear-with-bias weights are named through their
    bias, bias-free ones through the execution order of the architecture."""
    rng = np.random.default_rng(3)
    D = 6
    names = [f"dit.transformer_blocks.{i}.attn.{leaf}" for i in range(2) for leaf in ("to_q", "gate", "to_out.0")]
    W = {n: rng.standard_normal((D, D)).astype(np.float32) for n in names}
    inits, nodes = [], []
    for i in range(2):
        sfx = "" if i == 0 else f"_{i}"
        p = f"dit.transformer_blocks.{i}.attn."
        inits += [_tensor(f"onnx::MatMul_{10 * i + k}", [D, D], 1, raw=np.ascontiguousarray(W[p + leaf].T).tobytes())
                  for k, leaf in enumerate(("to_q", "gate", "to_out.0"))]
        inits.append(_tensor(f"m.{p}to_q.bias", [D], 1, raw=np.zeros(D, np.float32).tobytes()))
        nodes += [
            _node(f"/to_q{sfx}/MatMul", "MatMul", [f"x{i}", f"onnx::MatMul_{10 * i}"], [f"q{i}"]),
            _node(f"/to_q{sfx}/Add", "Add", [f"m.{p}to_q.bias", f"q{i}"], [f"qb{i}"]),
            _node(f"/gate{sfx}/MatMul", "MatMul", [f"x{i}", f"onnx::MatMul_{10 * i + 1}"], [f"g{i}"]),
            _node(f"/to_out/to_out.0{sfx}/MatMul", "MatMul", [f"g{i}", f"onnx::MatMul_{10 * i + 2}"], [f"x{i + 1}"]),
        ]
    path = tmp_path / "d.onnx"
    path.write_bytes(_model(nodes, inits))
    # declaration order puts to_out before gate on purpose: only the execution ranking gets it right
    specs = [(f"dit.transformer_blocks.{i}.attn.{leaf}.weight", (D, D)) for i in range(2)
             for leaf in ("to_q", "to_out.0", "gate")] + [(f"dit.transformer_blocks.{i}.attn.to_q.bias", (D,)) for i in range(2)]
    sd = weights.load_model_weights([str(path)], specs, "toy", exec_rank=weights.dit_exec_rank)
    for n in names:
        np.testing.assert_array_equal(np.asarray(sd[n + ".weight"]), W[n], err_msg=n)


def test_folded_log_scale_is_recovered_as_a_logarithm(tmp_path):
    """style.py:167: x * exp(log_scale).  Constant folding leaves exp(log_scale) as an anonymous scalar Mul operand in
    the style encoder's scope; an unfolded graph keeps the named parameter.  Both must load as log_scale = -1.8."""
    folded = _model(
        [_node("/style_encoder/Mul", "Mul", ["h", "onnx::Mul_7"], ["hs"])],
        [_tensor("onnx::Mul_7", [], 1, raw=np.array(np.exp(-1.8), "<f4").tobytes())])
    named = _model(
        [_node("/style_encoder/Exp", "Exp", ["m.style_encoder.log_scale"], ["s"]),
         _node("/style_encoder/Mul", "Mul", ["h", "s"], ["hs"])],
        [_tensor("m.style_encoder.log_scale", [], 1, raw=np.array(-1.8, "<f4").tobytes())])
    specs = [("style_encoder.log_scale", ())]
    for name, blob in (("folded", folded), ("named", named)):
        path = tmp_path / f"{name}.onnx"
        path.write_bytes(blob)
        sd = weights.load_model_weights([str(path)], specs, name)
        assert np.asarray(sd["style_encoder.log_scale"]).shape == ()
        assert abs(float(sd["style_encoder.log_scale"]) + 1.8) < 1e-6, name


def test_dit_exec_rank_follows_the_reference_forward():
    r = weights.dit_exec_rank
    assert r("style_encoder.blocks.11.mlp.w2.weight") < r("phoneme_embedding.blocks.0.attention.wq.weight")
    assert r("phoneme_embedding.blocks.7.mlp.w2.weight") < r("dit.transformer_blocks.0.attn.gate.weight")
    blk = "style_encoder.blocks.3."
    order = ["attention.wq", "attention.wk", "attention.wv", "attention.gate", "attention.wo", "mlp.w1", "mlp.w3", "mlp.w2"]
    assert sorted(order, key=lambda x: r(blk + x + ".weight")) == order  # style.py:47-66,76-77
    d = "dit.transformer_blocks.5."
    assert r(d + "attn.gate.weight") < r(d + "attn.to_out.0.weight") < r(d + "ff.w1.weight")  # dit.py:110-119
    assert r(d + "ff.w2.weight") < r("dit.transformer_blocks.6.attn.to_q.weight")


def test_onnx_reader_rejects_garbage(tmp_path):
    p = tmp_path / "bad.onnx"
    p.write_bytes(b"\x0a\xff\xff\xff\x0f" + b"x" * 10)  # length-delimited field longer than the file
    with pytest.raises(ValueError):
        weights.read_onnx(str(p))
    p.write_bytes(_v(1, 8))
    with pytest.raises(ValueError, match="no graph"):
        weights.read_onnx(str(p))


def test_onnx_export_of_hf_decoder_fixture():
    """tests/golden/tiny_codec_decoder.onnx is a torch.onnx (TorchScript exporter) export of a scaled-down HF
    VibeVoice decoder (oracle/make_golden_onnx.py): named Conv weights, anonymous transposed MatMul weights and
    constant-folded layer scales must all come back under their state-dict names, bit-exactly."""
    ref = np.load(os.path.join(GOLDEN, "tiny_codec_decoder.npz"))
    specs = [(k, ref[k].shape) for k in ref.files]
    sd = weights.load_model_weights([os.path.join(GOLDEN, "tiny_codec_decoder.onnx")], specs, "tiny decoder")
    assert sorted(sd) == sorted(ref.files)
    for k in ref.files:
        np.testing.assert_array_equal(np.asarray(sd[k]), ref[k], err_msg=k)
    # without the architecture list the folded layer scales cannot be named: strict matching says which
    bare = weights.onnx_state_dict(os.path.join(GOLDEN, "tiny_codec_decoder.onnx"))
    with pytest.raises(KeyError, match="gamma"):
        weights.match_to_specs(bare, specs, "tiny decoder")


def test_pt_checkpoint_prefixes_and_containers(tmp_path):
    """scripts/train/dmd2/distill.py:39-57,468-479: EMA / DDP / torch.compile prefixes, 'initted'/'step' extras."""
    sd = {"a.weight": torch.randn(3, 2), "b": torch.randn(4)}
    wrapped = {"ema_model.module._orig_mod." + k: v for k, v in sd.items()}
    wrapped.update({"initted": torch.tensor(True), "step": torch.tensor(7)})
    p = tmp_path / "ck.pt"
    torch.save({"student_model": wrapped, "optimizer": {"lr": 1.0}}, p)
    got = weights.load_state_dict_file(str(p))
    assert sorted(got) == ["a.weight", "b"]
    assert torch.equal(got["a.weight"], sd["a.weight"])
    torch.save(sd, p)
    assert sorted(weights.load_state_dict_file(str(p))) == ["a.weight", "b"]
    with pytest.raises(FileNotFoundError):
        weights.load_state_dict_file(str(tmp_path / "nope.pt"))


def test_safetensors_and_packed_roundtrip(tmp_path):
    from safetensors.numpy import save_file

    rng = np.random.default_rng(1)
    sd = {"x.weight": rng.standard_normal((8, 5)).astype(np.float32), "x.bias": rng.standard_normal(8).astype(np.float32),
          "scale": np.float32(-1.8).reshape(())}
    save_file({k: np.ascontiguousarray(v) for k, v in sd.items()}, str(tmp_path / "m.safetensors"))
    got = weights.load_state_dict_file(str(tmp_path / "m.safetensors"))
    for k in sd:
        np.testing.assert_array_equal(got[k], sd[k])

    weights.save_packed(str(tmp_path / "m.sttsw"), sd)
    got = weights.load_state_dict_file(str(tmp_path / "m.sttsw"))
    for k in sd:
        assert got[k].shape == sd[k].shape
        np.testing.assert_array_equal(got[k], sd[k])

    weights.save_packed(str(tmp_path / "h.sttsw"), sd, dtype="bfloat16")
    got = weights.load_packed(str(tmp_path / "h.sttsw"))
    np.testing.assert_array_equal(got["x.bias"], sd["x.bias"])  # 1-D tensors stay fp32
    want = torch.from_numpy(sd["x.weight"]).to(torch.bfloat16).float().numpy()  # round-to-nearest-even
    np.testing.assert_array_equal(got["x.weight"], want)
    assert os.path.getsize(tmp_path / "h.sttsw") < os.path.getsize(tmp_path / "m.sttsw")
    (tmp_path / "bad.sttsw").write_bytes(b"NOTMAGIC" + struct.pack("<Q", 0))
    with pytest.raises(ValueError):
        weights.load_packed(str(tmp_path / "bad.sttsw"))


class _ShapeOnly:
    def __init__(self, shape):
        self.shape = tuple(shape)


def test_match_to_specs_reports_by_name():
    specs = [("blk.w", (4, 3)), ("blk.b", (4,)), ("g", ())]
    sd = {"model.blk.w": np.zeros((3, 4), np.float32), "model.blk.b": np.zeros((4,), np.float32), "g": np.zeros(())}
    out = weights.match_to_specs(sd, specs, "toy")
    assert out["blk.w"].shape == (4, 3)  # transposed back
    with pytest.raises(KeyError, match="blk.b"):
        weights.match_to_specs({k: v for k, v in sd.items() if not k.endswith("blk.b")}, specs, "toy")
    with pytest.raises(KeyError, match="mis-shaped"):
        weights.match_to_specs({**sd, "model.blk.b": np.zeros((5,), np.float32)}, specs, "toy")
    with pytest.raises(KeyError, match="ambiguous"):
        weights.match_to_specs({**sd, "other.blk.b": np.zeros((4,), np.float32)}, specs, "toy")


def test_vibevoice_native_names_map_onto_the_hf_architecture():
    """microsoft/VibeVoice's own module tree (upsample_layers / downsample_layers / stages, SConv1d wrappers) ->
    HF names: the full decoder and encoder tensor lists must be covered exactly."""
    import re

    def to_native(name: str, up: bool) -> str:
        layers = "upsample_layers" if up else "downsample_layers"
        name = re.sub(r"^stem\.stage\.(\d+)\.", r"stages.0.\1.", name)
        name = re.sub(r"^stem\.", layers + ".0.0.", name)
        name = re.sub(r"^conv_layers\.(\d+)\.stage\.(\d+)\.", lambda m: f"stages.{int(m.group(1)) + 1}.{m.group(2)}.", name)
        name = re.sub(r"^conv_layers\.(\d+)\.", lambda m: f"{layers}.{int(m.group(1)) + 1}.0.", name)
        return name.replace("mixer.conv.", "mixer.conv.conv.conv.").replace("head.conv.", "head.conv.conv.")

    for specs, up in ((synthetic.vocoder_specs(), True), (synthetic.encoder_specs(), False)):
        native = {"model.acoustic_tokenizer.x." + to_native(n, up): _ShapeOnly(s) for n, s, *_ in specs}
        assert not any(k.endswith(n) for k in native for n, *_ in specs if ".stage." in n)
        out = weights.match_to_specs(weights.vibevoice_native_to_hf(native), specs, "codec")
        assert len(out) == len(specs)


def test_spec_tables_match_survey_counts():
    n = lambda specs: sum(int(np.prod(s[1], dtype=np.int64)) for s in specs)  # noqa: E731
    assert n(synthetic.dit_specs()) == 327_756_609
    assert n(synthetic.vocoder_specs()) == 343_695_969
    assert n(synthetic.encoder_specs()) == 343_696_032


def test_convert_weights_cli_roundtrip(tmp_path, monkeypatch):
    """tools/convert_weights.py: an exported .onnx -> .sttsw (bf16 matrices) -> load_model_weights again."""
    import importlib.util

    spec = importlib.util.spec_from_file_location(
        "convert_weights", os.path.join(os.path.dirname(__file__), "..", "tools", "convert_weights.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ref = np.load(os.path.join(GOLDEN, "tiny_codec_decoder.npz"))
    specs = [(k, ref[k].shape) for k in ref.files]
    monkeypatch.setattr(synthetic, "vocoder_specs", lambda: specs)
    out = str(tmp_path / "dec.sttsw")
    assert mod.main(["decoder", os.path.join(GOLDEN, "tiny_codec_decoder.onnx"), "-o", out, "--bf16"]) == 0
    sd = weights.load_model_weights([out], specs, "tiny decoder")
    for k in ref.files:
        want = ref[k] if ref[k].ndim < 2 else torch.from_numpy(ref[k]).to(torch.bfloat16).float().numpy()
        np.testing.assert_array_equal(np.asarray(sd[k]), want, err_msg=k)

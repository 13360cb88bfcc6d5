"""CPU-side checks: the C-ABI library loads, exports every symbol include/smalltts_b200.h declares, and the host
logic (padding, checkpoint reading, duration rules) behaves like the reference.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from smalltts_b200 import build, _cabi

    build.build()
    return _cabi.lib()


def test_header_symbols_are_exported_and_bound(lib):
    from smalltts_b200 import _cabi

    header = open(os.path.join(ROOT, "include", "smalltts_b200.h")).read()
    declared = set(re.findall(r"\b(stts_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_cabi.SIGNATURES), declared ^ set(_cabi.SIGNATURES)


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_engine_fails_loudly_without_gpu(lib):
    from smalltts_b200.engine import Engine

    with pytest.raises(RuntimeError, match="no CPU fallback|no CUDA device"):
        Engine(0)


def test_missing_library_is_an_error(monkeypatch, tmp_path):
    from smalltts_b200 import _cabi

    monkeypatch.setattr(_cabi, "_libs", {})
    monkeypatch.setattr(_cabi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _cabi.lib()
    with pytest.raises(ValueError, match="precision"):
        _cabi.lib("exact")


def test_parity_build_exports_the_same_abi(lib):
    """precision="tight": the same sources compiled with fp16 instead of bf16 operands (csrc/op16.cuh) -- a second
    library with the identical C ABI."""
    from smalltts_b200 import _cabi, build

    build.build(tight=True)
    tight = _cabi.lib("tight")
    assert tight is not lib
    for name in _cabi.SIGNATURES:
        assert hasattr(tight, name), name


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """The ctypes mirrors in _cabi.py must have the size and field offsets the C compiler gives the structs of
    include/smalltts_b200.h (a drifted test-hook struct would corrupt arguments silently)."""
    import subprocess

    from smalltts_b200 import _cabi

    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stddef.h>\n#include <stdio.h>\n#include "smalltts_b200.h"\n'
        "int main(void) {\n"
        '  printf("config %zu %zu\\n", sizeof(stts_config), offsetof(stts_config, reserved));\n'
        '  printf("timing %zu %zu\\n", sizeof(stts_timing), offsetof(stts_timing, total_ms));\n'
        '  printf("chain %zu %zu %zu %zu %zu %zu\\n", sizeof(stts_test_chain_args), offsetof(stts_test_chain_args, ready),\n'
        "         offsetof(stts_test_chain_args, kv_ref), offsetof(stts_test_chain_args, M), offsetof(stts_test_chain_args, qkv_db),\n"
        "         offsetof(stts_test_chain_args, blk));\n"
        "  return 0;\n}\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict((l.split()[0], [int(x) for x in l.split()[1:]]) for l in subprocess.run([str(exe)], check=True, capture_output=True,
                                                                                      text=True).stdout.splitlines())
    assert out["config"] == [ctypes.sizeof(_cabi.Config), _cabi.Config.reserved.offset]
    assert out["timing"] == [ctypes.sizeof(_cabi.Timing), _cabi.Timing.total_ms.offset]
    ca = _cabi.ChainArgs
    assert out["chain"] == [ctypes.sizeof(ca), ca.ready.offset, ca.kv_ref.offset, ca.M.offset, ca.qkv_db.offset, ca.blk.offset]


def test_duration_and_frame_rules_match_reference():
    from smalltts_b200 import infer

    assert infer.frames_for(10.0) == 75 and infer.frames_for(2.0) == 15 and infer.frames_for(5.0) == 37  # floor
    assert infer.frames_for(0.01) == 1
    assert infer.estimate_duration("a" * 1000) == 30.0


def test_pad_batch():
    from smalltts_b200.engine import pad_batch

    ref, rl, ids, pl = pad_batch([np.ones((3, 64), np.float32), torch.ones(5, 64)], [[1, 2], [3, 4, 5, 6]], [4, 4])
    assert ref.shape == (2, 5, 64) and rl == [3, 5] and ids.shape == (2, 4) and pl == [2, 4]
    assert ref[0, 3:].sum() == 0 and ids[0, 2:].sum() == 0
    with pytest.raises(ValueError):
        pad_batch([np.ones((3, 32), np.float32)], [[1]], [1])


def test_checkpoint_prefix_stripping(tmp_path):
    from smalltts_b200.infer import load_state_dict_file

    sd = {"ema_model.module.velocity.bias": torch.zeros(64), "initted": torch.tensor(True), "step": torch.tensor(3)}
    p = tmp_path / "ckpt.pt"
    torch.save({"student_model": sd}, p)
    out = load_state_dict_file(str(p))
    assert list(out) == ["velocity.bias"]
    with pytest.raises(FileNotFoundError):
        load_state_dict_file(str(tmp_path / "missing.pt"))


def test_synthetic_weights_have_reference_shapes():
    from smalltts_b200 import synthetic

    assert sum(int(np.prod(s)) for _, s, _, _ in synthetic.dit_specs()) == 327_756_609
    assert sum(int(np.prod(s)) for _, s, _, _ in synthetic.vocoder_specs()) == 343_695_969
    assert len(synthetic.dit_specs()) == 592 and len(synthetic.vocoder_specs()) == 276
    assert sum(int(np.prod(s)) for _, s, _, _ in synthetic.encoder_specs()) == 343_696_032  # hf encoder


def test_ensure_assets_offline_behaviour(tmp_path, monkeypatch):
    """assets/ensure.py:20-41 mirror: an existing folder is left alone; a missing one without a reachable Hub is a
    FileNotFoundError that names the files to provide (there is no network in the build / bench environment)."""
    import sys
    import types

    from smalltts_b200.assets import ensure_assets

    (tmp_path / "dmd").mkdir()
    ensure_assets(["dmd", ""], root=str(tmp_path))  # nothing to do
    fake = types.ModuleType("huggingface_hub")

    def snapshot_download(**kw):
        raise OSError("network unreachable")

    fake.snapshot_download = snapshot_download
    monkeypatch.setitem(sys.modules, "huggingface_hub", fake)
    with pytest.raises(FileNotFoundError, match="codec/decoder.onnx"):
        ensure_assets("codec", root=str(tmp_path))
    calls = []
    fake.snapshot_download = lambda **kw: calls.append(kw) or (tmp_path / "codec").mkdir()
    ensure_assets("codec", root=str(tmp_path))
    assert calls[0]["repo_id"] == "smallbraineng/smalltts" and calls[0]["allow_patterns"] == ["codec/*"]

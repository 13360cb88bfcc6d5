"""The C++ host side above the C ABI (include/smalltts_b200_pipeline.hpp, the mirror of the reference's Rust
`Pipeline`, src/server/src/pipeline.rs): it compiles and links against libsmalltts_b200.so with plain g++, reads the
`.sttsw` containers exactly like the Python reader, and fails loudly (no fallback) without a CUDA device."""
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

from smalltts_b200 import _cabi, weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GXX = shutil.which("g++")
pytestmark = pytest.mark.skipif(GXX is None or not os.path.exists(_cabi.LIB_PATH), reason="needs g++ and the built library")


def _compile(src, out):
    libdir = os.path.dirname(_cabi.LIB_PATH)
    cmd = [GXX, "-O1", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-L", libdir,
           "-lsmalltts_b200", f"-Wl,-rpath,{libdir}", "-o", out]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    return out


def test_sttsw_reader_matches_python(tmp_path):
    rng = np.random.default_rng(0)
    sd = {"blk.0.weight": rng.standard_normal((6, 5)).astype(np.float32), "blk.0.bias": rng.standard_normal(6).astype(np.float32),
          "style_encoder.log_scale": np.float32(-1.8).reshape(()), "conv.weight": rng.standard_normal((4, 3, 7)).astype(np.float32)}
    exe = _compile(os.path.join(ROOT, "tests", "cpp", "read_sttsw_main.cpp"), str(tmp_path / "read_sttsw"))
    for dtype in ("float32", "bfloat16"):
        path = str(tmp_path / f"w_{dtype}.sttsw")
        weights.save_packed(path, sd, dtype=dtype)
        want = weights.load_packed(path)
        r = subprocess.run([exe, path], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert r.returncode == 0, r.stderr
        rows = [json.loads(line) for line in r.stdout.splitlines()]
        assert [row["name"] for row in rows] == list(sd)
        for row in rows:
            w = np.asarray(want[row["name"]], dtype=np.float32)
            assert tuple(row["shape"]) == w.shape
            assert abs(row["sum"] - float(w.astype(np.float64).sum())) <= 1e-5
            assert np.float32(row["first"]) == w.reshape(-1)[0] and np.float32(row["last"]) == w.reshape(-1)[-1]
    bad = tmp_path / "bad.sttsw"
    bad.write_bytes(b"NOTMAGIC" + b"\0" * 16)
    r = subprocess.run([exe, str(bad)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "not a .sttsw file" in r.stderr


def test_pipeline_example_builds_links_and_has_no_cpu_fallback(tmp_path):
    import torch

    exe = _compile(os.path.join(ROOT, "examples", "bench_pipeline.cpp"), str(tmp_path / "bench_pipeline"))
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 2 and "usage" in r.stderr
    if torch.cuda.is_available():
        pytest.skip("the no-device error path needs a machine without a GPU")
    r = subprocess.run([exe, "a.sttsw", "b.sttsw", "c.sttsw"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr

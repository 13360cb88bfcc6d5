"""The C++ host side (include/smalltts_b200_pipeline.hpp, mirror of the reference's Rust Pipeline, pipeline.rs:40-112)
EXECUTED on a B200: compiled with g++, linked against libsmalltts_b200.so, fed .sttsw weight files, compared with the
Python serving shim on the same requests and seed."""
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_cpp_pipeline_runs_and_matches_the_python_pipeline(tmp_path):
    from smalltts_b200 import _cabi, serve, synthetic, weights
    from smalltts_b200.infer import SmallTTS

    sds = (synthetic.dit_state_dict(0), synthetic.vocoder_state_dict(1), synthetic.encoder_state_dict(2))
    paths = []
    for name, sd in zip(("dit", "decoder", "encoder"), sds):
        paths.append(str(tmp_path / f"{name}.sttsw"))
        weights.save_packed(paths[-1], sd)
    libdir = os.path.dirname(_cabi.LIB_PATH)
    exe = str(tmp_path / "pipeline_main")
    cmd = ["g++", "-O1", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "pipeline_main.cpp"), "-L", libdir, "-lsmalltts_b200", f"-Wl,-rpath,{libdir}", "-o", exe]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    out = str(tmp_path / "audio.f32")
    r = subprocess.run([exe, *paths, out], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert (info["n0"], info["n1"], info["n_single"]) == (8 * 3200, 4 * 3200, 15 * 3200)  # ceil(duration * 7.5) frames
    assert info["denoise_ms"] > 0 and info["codec_enc_ms"] > 0 and info["batch_total_ms"] > info["denoise_ms"]
    assert info["repeat_abs_diff"] > 0  # a fresh noise stream per request without an explicit seed
    got = np.fromfile(out, dtype=np.float32)
    assert got.size == 12 * 3200 and np.isfinite(got).all() and got.std() > 0

    # the same two requests through the Python shim with the same seed: the same engine pass, hence the same audio
    t = np.arange(int(2.0 * 24000), dtype=np.float32)
    refs = [(0.3 * np.sin(2 * np.float32(3.14159265358979) * 440.0 * t / 24000.0)).astype(np.float32),
            (0.3 * np.sin(2 * np.float32(3.14159265358979) * 220.0 * t[: int(1.2 * 24000)] / 24000.0)).astype(np.float32)]
    tts = SmallTTS(state_dicts=sds, seed=1234)
    try:
        want, _ = serve.Pipeline(tts).synthesize_many(refs, [[5, 9, 20, 33, 7], [101, 3, 44]], [1.01, 0.5])
    finally:
        tts.engine.close()
    want = np.concatenate(want)
    err = float(np.linalg.norm(got - want) / np.linalg.norm(want))
    print("C++ vs Python pipeline rel-L2", err)
    assert err <= 2e-2  # the sine is computed in float by both, to within an ulp: inputs differ by ~1e-7

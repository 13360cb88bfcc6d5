"""ctypes binding of include/smalltts_b200.h.  No torch types cross this boundary.

The shared library is built in-tree by ``python -m smalltts_b200.build``.  If it is missing this module
raises: there is deliberately no Python/CPU fallback for the hot path.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# STTS_LIB_PATH: load another build of the library (A/B experiments with compile-time variants, tools/build_variant.py)
LIB_PATH = os.environ.get("STTS_LIB_PATH") or os.path.join(HERE, "libsmalltts_b200.so")

MEM_HOST, MEM_DEVICE = 0, 1
OK = 0

c_i64p = C.POINTER(C.c_int64)
c_i32p = C.POINTER(C.c_int32)
c_f32p = C.POINTER(C.c_float)
vp = C.c_void_p


class Config(C.Structure):
    _fields_ = [("device", C.c_int), ("reserved", C.c_int * 7)]


class Timing(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("codec_enc_ms", "cond_enc_ms", "denoise_ms", "codec_dec_ms", "total_ms")]


class ChainArgs(C.Structure):
    """stts_test_chain_args (include/smalltts_b200.h)."""

    _fields_ = ([(n, vp) for n in ("wqkvg", "wo", "w13", "w2", "wvel", "bqkvg", "b13", "b2", "bvel", "qn", "kn", "cos_t",
                                   "sin_t", "x", "xb", "stats", "qkv", "gate", "ob", "hb", "vel", "ready", "frames", "mod",
                                   "fold", "trace", "kv_ref", "kv_text", "ref_len", "ph_len")]
                + [(n, C.c_int32) for n in ("M", "T", "n_phases", "B", "R", "P", "qkv_db")]
                + [("kind", C.c_int32 * 64), ("blk", C.c_int32 * 64)])


# name -> (restype, argtypes); kept in one table so tests can check every symbol of the header is exported
SIGNATURES = {
    "stts_create": (C.c_int, [C.POINTER(Config), C.POINTER(vp)]),
    "stts_destroy": (None, [vp]),
    "stts_engine_clone": (C.c_int, [vp, C.POINTER(vp)]),
    "stts_last_error": (C.c_char_p, [vp]),
    "stts_load_weight": (C.c_int, [vp, C.c_int, C.c_char_p, vp, C.c_int, c_i64p]),
    "stts_finalize_weights": (C.c_int, [vp]),
    "stts_encode_conditions": (C.c_int, [vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]),
    "stts_cond_free": (None, [vp, vp]),
    "stts_cond_read_kv": (C.c_int, [vp, vp, C.c_int, C.c_int, vp]),
    "stts_denoise_step": (C.c_int, [vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp]),
    "stts_sample": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_uint64, C.c_int, vp]),
    "stts_sample_teacher": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, vp, C.c_uint64,
                                      C.c_int, vp]),
    "stts_decode": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, vp]),
    "stts_encode_audio": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, vp]),
    "stts_resample_length": (C.c_int64, [C.c_int, C.c_int, C.c_int]),
    "stts_resample": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "stts_synthesize": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp,
                                  C.c_uint64, C.c_int, vp]),
    "stts_get_timings": (C.c_int, [vp, C.POINTER(Timing)]),
    "stts_launch_count": (C.c_uint64, []),
    "stts_last_vocoder_ms": (C.c_float, [vp, C.c_int]),
    "stts_timer_start": (C.c_int, [vp]),
    "stts_timer_stop": (C.c_int, [vp, c_f32p]),
    "stts_host_alloc": (vp, [C.c_size_t]),
    "stts_host_free": (None, [vp]),
    "stts_test_set_async": (C.c_int, [vp, C.c_int]),
    "stts_test_gemm": (C.c_int, [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int,
                                 vp, C.c_int, C.c_int, vp, vp, C.c_int, vp, C.c_int, vp, vp, C.c_int]),
    "stts_test_gemm_split": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, vp, C.c_int, vp, C.c_int, vp, vp, vp, vp]),
    "stts_test_attention": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, C.c_int, vp, vp,
                                      vp, C.c_int, vp, C.c_int, vp]),
    "stts_test_convnext_mix": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp]),
    "stts_test_ffn_fused": (C.c_int, [vp, vp, vp, C.c_longlong, C.c_int, vp, vp, vp, vp, vp, vp, vp]),
    "stts_test_chain": (C.c_int, [vp, C.POINTER(ChainArgs)]),
    "stts_test_chain_fold": (C.c_int, [vp, C.POINTER(ChainArgs), vp]),
    "stts_test_chain_fold_floats": (C.c_int64, []),
    "stts_test_chain_stats_cast": (C.c_int, [vp, vp, C.c_int, vp, vp, vp]),
    "stts_test_convnext_fused": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
}

# "tight": the parity build of the same sources with fp16 (11-bit significand, TF32's precision) instead of bf16 GEMM /
# attention operands (csrc/op16.cuh).  Same C ABI, same kernels; an engine picks its library when it is created.
LIB_PATH_TIGHT = os.environ.get("STTS_LIB_PATH_TIGHT") or os.path.join(HERE, "libsmalltts_b200_tight.so")
PRECISIONS = {"fast": "bf16 operands, fp32 accumulation", "tight": "fp16 operands (11-bit significand), fp32 accumulation"}

_libs = {}


def lib(precision: str = "fast") -> C.CDLL:
    if precision not in PRECISIONS:
        raise ValueError(f"precision must be one of {sorted(PRECISIONS)}, got {precision!r}")
    if precision not in _libs:
        path = LIB_PATH_TIGHT if precision == "tight" else LIB_PATH
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} is missing: build it with `python -m smalltts_b200.build`. "
                "smalltts_b200 has no CPU or PyTorch fallback for the synthesize path."
            )
        l = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _libs[precision] = l
    return _libs[precision]


def check(rc: int, handle=None, precision: str = "fast") -> None:
    if rc == OK:
        return
    msg = lib(precision).stts_last_error(handle)
    msg = msg.decode() if msg else "unknown error"
    if rc == -1:
        raise ValueError(f"smalltts_b200: {msg}")
    raise RuntimeError(f"smalltts_b200 (status {rc}): {msg}")

#!/usr/bin/env python
"""Micro-benchmark of the tcgen05 GEMM on the engine's hot-path shapes (CUDA events on the engine stream).

usage: bench_gemm.py [reps] > gpurun_out/gemm_bench.txt
For every shape x tile width: warm (same operands every launch) and cold-W (weights rotated through > L2 worth of
copies, the steady state of the DiT where 410 MB of weights stream from HBM every step)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from smalltts_b200 import _cabi
from smalltts_b200.engine import Engine

PAIR = 0x1000  # gemm.cuh kGemmPairFlag: CTA-pair variant (cta_group::2), printed as bn + 4096
ACT = dict(none=0, gelu=1, mish=2, swiglu=4, gelu2=1 + 16, none_f16=0 + 32)


def p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    only = sys.argv[2] if len(sys.argv) > 2 else ""
    eng = Engine(0)
    lib = _cabi.lib()
    lib.stts_test_set_async(eng._h, 1)
    shapes = [
        # name, M, N, K, act, residual, out (f32|bf16)
        ("dit_qkvg", 600, 3840, 960, "none", False, "f32"),
        ("dit_wo", 600, 960, 1024, "none", True, "f32"),
        ("dit_w13", 600, 4800, 960, "swiglu", False, "bf16"),
        ("dit_w2", 600, 960, 2400, "none", True, "f32"),
        ("stem_1", 600, 8192, 2048, "gelu2", False, "bf16"),
        ("stem_2", 600, 2048, 8192, "none_f16", True, "f32"),
        ("up0_1", 4800, 4096, 1024, "gelu2", False, "bf16"),
        ("up0_2", 4800, 1024, 4096, "none_f16", True, "f32"),
        ("up1_1", 24000, 2048, 512, "gelu2", False, "bf16"),
        ("up1_2", 24000, 512, 2048, "none_f16", True, "f32"),
        ("up2_1", 120000, 1024, 256, "gelu2", False, "bf16"),
        ("up2_2", 120000, 256, 1024, "none_f16", True, "f32"),
        ("up3_1", 480000, 512, 128, "gelu2", False, "bf16"),
        ("up3_2", 480000, 128, 512, "none_f16", True, "f32"),
    ]
    print(f"{'shape':10s} {'M':>7s} {'N':>5s} {'K':>5s} {'bn':>4s} {'warm us':>9s} {'TF/s':>7s} {'coldW us':>9s} {'TF/s':>7s}")
    for name, M, N, K, act, res, out in shapes:
        if only and not any(o in name for o in only.split(",")):
            continue
        torch.manual_seed(0)
        a = (torch.randn(M, K, device="cuda")).to(torch.bfloat16)
        wbytes = N * K * 2
        ncopies = max(1, min(48, (260 << 20) // wbytes))
        ws = [(torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16) for _ in range(ncopies)]
        bias = torch.randn(N, device="cuda")
        ncol = N // 2 if act == "swiglu" else N
        o32 = torch.zeros(M, ncol, device="cuda") if out == "f32" else None
        o16 = torch.zeros(M, ncol, device="cuda", dtype=torch.bfloat16) if out == "bf16" else None
        r = torch.randn(M, ncol, device="cuda") if res else None
        flops = 2.0 * M * N * K
        torch.cuda.synchronize()  # the engine stream does not wait for torch's
        for bn in (32, 64, 128, 256, 128 | PAIR, 256 | PAIR):
            if (bn & ~PAIR) > N or (bn == 32 and M > 5000):
                continue
            if bn & PAIR and act not in ("none", "none_f16", "gelu2"):
                continue  # pair instantiations: ACT_NONE and the fp16 2*gelu epilogue

            def run(w):
                rc = lib.stts_test_gemm(eng._h, bn, p(a), 1, M, K, K, p(w), N, K, N, K, 1, 0, 1, 1, 0, 0, 0, p(bias),
                                        ACT[act], None, 0, 0, None, None, 0, p(r), ncol if res else 0, p(o32), p(o16), ncol)
                _cabi.check(rc, eng._h)

            res_us = []
            for mode in ("warm", "cold"):
                for _ in range(3):
                    run(ws[0])
                ms = C.c_float()
                _cabi.check(lib.stts_timer_start(eng._h), eng._h)
                for i in range(reps):
                    run(ws[i % ncopies] if mode == "cold" else ws[0])
                _cabi.check(lib.stts_timer_stop(eng._h, C.byref(ms)), eng._h)
                res_us.append(ms.value * 1e3 / reps)
            print(f"{name:10s} {M:7d} {N:5d} {K:5d} {bn:4d} {res_us[0]:9.2f} {flops / res_us[0] / 1e6:7.1f} "
                  f"{res_us[1]:9.2f} {flops / res_us[1] / 1e6:7.1f}", flush=True)
        del ws, a
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

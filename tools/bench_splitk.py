#!/usr/bin/env python
"""Split-K GEMM micro-benchmark (stts_test_gemm_split in async mode times 20 launches on the engine stream)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smalltts_b200 import _cabi
from smalltts_b200.engine import Engine

eng = Engine(0)
lib = _cabi.lib()
lib.stts_test_set_async(eng._h, 1)
p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
for (M, N, K, bn, gelu2) in [(600, 8192, 2048, 256, True), (600, 2048, 8192, 128, False), (600, 8192, 4096, 256, False), (4800, 1024, 4096, 256, False)]:
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    o32 = torch.zeros(M, N, device="cuda")
    o16 = torch.zeros(M, N, device="cuda", dtype=torch.float16)
    torch.cuda.synchronize()
    row = []
    for S in (1, 2, 3, 4):
        rc = lib.stts_test_gemm_split(eng._h, bn, S, p(a), M, K, p(w), N, p(bias), 1 if gelu2 else 0, None, None, None if gelu2 else p(o32), p(o16) if gelu2 else None)
        _cabi.check(rc, eng._h)
        row.append(round(lib.stts_last_vocoder_ms(eng._h, 0) * 1e3, 1))
    print(f"M={M} N={N} K={K} bn={bn} gelu2={gelu2}: us per launch for splits 1..4 = {row}")

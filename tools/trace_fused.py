#!/usr/bin/env python
"""Role timeline of the fused ConvNeXt tail kernel (CTA 0), from a debug build:
    STTS_EXTRA_NVCC_FLAGS=-DSTTS_FUSED_TRACE python -m smalltts_b200.build --force && python tools/trace_fused.py [C]
Prints clock64() deltas (cycles) of the role events of tiles 8..15, relative to each tile's 'data landed' event."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from smalltts_b200 import _cabi
from smalltts_b200.engine import Engine

Cc = int(sys.argv[1]) if len(sys.argv) > 1 else 64
eng = Engine(0)
lib = _cabi.lib()
B, T = 8, 120000 if Cc == 64 else 240000
x = torch.randn(B, T, Cc, device="cuda")
p = lambda t: C.c_void_p(t.data_ptr())
nw, fw = torch.ones(Cc, device="cuda"), torch.ones(Cc, device="cuda")
cw, cb = torch.randn(Cc, 7, device="cuda") * 0.3, torch.zeros(Cc, device="cuda")
g, fg = torch.full((Cc,), 0.1, device="cuda"), torch.full((Cc,), 0.1, device="cuda")
w1 = (torch.randn(4 * Cc, Cc, device="cuda") * Cc ** -0.5).to(torch.bfloat16)
w2 = (torch.randn(Cc, 4 * Cc, device="cuda") * (4 * Cc) ** -0.5 * 0.5).to(torch.float16)
b1, b2 = torch.zeros(4 * Cc, device="cuda"), torch.zeros(Cc, device="cuda")
out = torch.empty_like(x)
for _ in range(2):
    rc = lib.stts_test_convnext_fused(eng._h, p(x), B, T, Cc, p(nw), p(cw), p(cb), p(g), p(fw), p(w1), p(b1), p(w2), p(b2),
                                      p(fg), p(out), None)
    _cabi.check(rc, eng._h)
torch.cuda.synchronize()
buf = (C.c_longlong * (64 * 16))()
lib.stts_debug_fused_trace.restype = C.c_int
assert lib.stts_debug_fused_trace(buf) == 0
tr = np.array(buf[:], dtype=np.int64).reshape(64, 16)
names = ["mix:data", "mix:conv", "mix:a_full", "mma:mma1", "mma:o_commit", "gelu:h_full", "gelu:last", "out:y_copied",
         "out:o_full", "out:stored"]
t0 = tr[8, 0]
print("tile " + " ".join(f"{n:>13s}" for n in names))
for it in range(8, 20):
    print(f"{it:4d} " + " ".join(f"{tr[it, e] - t0:13d}" for e in range(10)))
print("mixer thread 0: tile | loop top | after try_prefetch | data seen | a_full | before blocking try | after")
for it in range(8, 14):
    print(f"   {it:3d} " + " ".join(f"{tr[it, e] - t0:10d}" for e in (10, 11, 0, 2, 12, 13)))
print("cycles per tile (data-landed deltas):", np.diff(tr[8:24, 0]).tolist())

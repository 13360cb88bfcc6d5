#!/usr/bin/env python
"""Role timeline of the chained DiT kernel (dit_chain.cu trace_ev) for one 4-phase launch at the benchmark shape
(M = 600 rows): where does a phase transition spend its time?

    python tools/trace_chain.py [split]      # 'split' = the same four GEMMs as four launches
    python tools/trace_chain.py attn         # a whole block with attention as a phase (M = 640 rows, 210 keys)

Events per tile (ns, globaltimer): 0 published | 1 A producer at the flag | 2 flag seen | 3 first operands landed |
4 MMAs issued | 5 epilogue sees the accumulator | 6 epilogue stores issued | 7 tile counted."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch

import test_gpu_chain as tc
from smalltts_b200 import _cabi
from smalltts_b200.engine import Engine

eng = Engine(0)
attn = len(sys.argv) > 1 and sys.argv[1] == "attn"
if attn:
    M, T = 640, 80
    c = tc.Case(M, T, [80] * 8, seed=3, R=75, P=60, ref_len=[75] * 8, ph_len=[55] * 8)
    c.qkv_db = 1
    phases = [(tc.QKVG, 1), (tc.ATTN, 1), (tc.OUT, 1), (tc.W13, 1), (tc.W2, 1), (tc.QKVG, 2), (tc.ATTN, 2)]
else:
    M, T = 600, 75
    c = tc.Case(M, T, [75] * 8, seed=3)
    phases = [(tc.OUT, 1), (tc.W13, 1), (tc.W2, 1), (tc.QKVG, 2)]
c.make_fold(eng)
m_tiles = (M + 127) // 128
NT = {tc.QKVG: 29, tc.OUT: 15, tc.W13: 25, tc.W2: 15, tc.VEL: 1}
ITEMS = {k: v * m_tiles for k, v in NT.items()}
ITEMS[tc.ATTN] = (M // T) * ((T + 127) // 128) * 8
NT[tc.ATTN] = 8
split = len(sys.argv) > 1 and sys.argv[1] == "split"
x0 = c.x.clone()
if attn:
    c.stats_cast(eng, c.m(1, 1))


def run(trace):
    c.x.copy_(x0)
    groups = [[p] for p in phases] if split else [phases]
    torch.cuda.synchronize()
    eng.timer_start()
    for k, g in enumerate(groups):
        a = c.args(g)
        a.trace = None if trace is None else C.c_void_p(trace[k].data_ptr())
        _cabi.check(_cabi.lib().stts_test_chain(eng._h, C.byref(a)), eng._h)
    return eng.timer_stop()


for _ in range(3):
    run(None)
ms = [run(None) for _ in range(5)]
print("mode", "split" if split else "fused", f"ms per chain ({len(phases)} phases):", [round(m, 4) for m in ms])
n_l = 4 if split else 1
trace = [torch.zeros(148 * 64 * 16, dtype=torch.int64, device="cuda") for _ in range(n_l)]
run(trace)
torch.cuda.synchronize()
t0 = None
rows = []
for k in range(n_l):
    tr = trace[k].cpu().numpy().reshape(148, 64, 16)
    ph_list = [phases[k]] if split else phases
    for cta in range(148):
        for s in range(64):
            g = int(tr[cta, s, 15])
            if g == 0:
                continue
            g -= 1
            p = 0
            while p + 1 < len(ph_list) and g >= ITEMS[ph_list[p][0]]:
                g -= ITEMS[ph_list[p][0]]
                p += 1
            rows.append((k if split else p, cta, s, g // NT[ph_list[p][0]], g % NT[ph_list[p][0]], tr[cta, s, :8].astype(np.int64),
                         tr[cta, s, 8:15].astype(np.int64)))
t0 = min(int(r[5][r[5] > 0].min()) for r in rows)
names = ["published", "A@flag", "flag seen", "operands", "MMAs issued", "epi sees acc", "epi stored", "counted"]
for p in range(len(phases)):
    ev = np.array([r[5] for r in rows if r[0] == p], dtype=np.float64)
    ev = np.where(ev > 0, (ev - t0) / 1e3, np.nan)
    print(f"--- phase {p} kind {phases[p][0]}: {ev.shape[0]} tiles; event times in us (min / median / max)")
    for i, n in enumerate(names):
        col = ev[:, i]
        print(f"   {n:14s} {np.nanmin(col):8.2f} {np.nanmedian(col):8.2f} {np.nanmax(col):8.2f}")
    d = lambda a, b: np.nanmedian(ev[:, b] - ev[:, a])  # noqa: E731
    if phases[p][0] == tc.ATTN:
        ia = np.array([r[6] for r in rows if r[0] == p], dtype=np.float64)
        ia = np.where(ia > 0, (ia - t0) / 1e3, np.nan)
        st = ev[:, 5:6]
        if os.environ.get("STTS_ATTN_DEBUG_TRACE"):
            print("   [debug build] P written per warp (half*4+q = 0..6), after the dependency:", np.round(np.nanmedian(ia - st, axis=0), 2),
                  "| MMA warp: V landed %.2f, P complete %.2f, PV issued %.2f" % tuple(np.nanmedian(ev[:, [3, 1, 4]] - st, axis=0)))
            print("   o_full seen %.2f | item done %.2f" % (np.nanmedian(ia[:, 6:7] - st), np.nanmedian(ev[:, 6:7] - st)))
            continue
        print("   item timeline after the dependency (median us): scores ready %.2f | P written %.2f | output ready %.2f | "
              "stored %.2f" % tuple(np.nanmedian(ia - st, axis=0)[:4]))
    print(f"   medians: flag wait {d(1, 2):.2f} | flag->operands {d(2, 3):.2f} | mainloop {d(3, 4):.2f} | "
          f"MMA tail {d(4, 5):.2f} | epilogue {d(5, 6):.2f} | publish {d(6, 7):.2f}")
per_cta = {}
for r in rows:
    per_cta.setdefault(r[1], []).append(r[0])
hist = {}
for v in per_cta.values():
    hist[tuple(sorted(v))] = hist.get(tuple(sorted(v)), 0) + 1
print("tiles per CTA (phase multiset -> CTAs):", sorted(hist.items(), key=lambda kv: -kv[1])[:12])

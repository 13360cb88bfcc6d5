#!/usr/bin/env python
"""Throughput of N engine replicas on ONE GPU (N host threads, N non-blocking streams): does a second in-flight batch
fill the SMs the latency-bound DiT GEMMs (75-150 tiles at M = 600) leave idle?
usage: bench_concurrent.py [steps] [max_replicas] [device]     (one human-readable line and one JSON line per count)"""
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from smalltts_b200 import synthetic
from smalltts_b200.engine import Engine, pad_batch

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
max_replicas = int(sys.argv[2]) if len(sys.argv) > 2 else 3
device = int(sys.argv[3]) if len(sys.argv) > 3 else 0
torch.cuda.set_device(device)
sds = (synthetic.dit_state_dict(0), synthetic.vocoder_state_dict(1))
refs, ids, frames, _ = synthetic.synthetic_inputs(8, 75, 15, 120)
ref, ref_len, idt, ph_len = pad_batch(refs, ids, frames)
engines = []
for n in range(1, max_replicas + 1):
    while len(engines) < n:
        if engines and os.environ.get("STTS_CLONE") == "1":
            e = engines[0][0].clone()  # shares the first engine's weights (stts_engine_clone)
        else:
            e = Engine(device)
            e.load_state_dicts(*sds)
        dev = [torch.from_numpy(ref).cuda(), torch.from_numpy(idt).cuda()]
        out = torch.empty(8, 75 * 3200, device="cuda")
        for i in range(3):
            e.synthesize(dev[0], ref_len, dev[1], ph_len, frames, 75, seed=i, out=out)
        engines.append((e, dev, out))
    torch.cuda.synchronize()

    def work(k):
        e, dev, out = engines[k]
        for i in range(steps):
            e.synthesize(dev[0], ref_len, dev[1], ph_len, frames, 75, seed=10 + i, out=out)

    dt = float("inf")
    for _ in range(2):  # wall clock around host threads: best of two passes (a host hiccup costs a whole pass otherwise)
        ths = [threading.Thread(target=work, args=(k,)) for k in range(n)]
        t0 = time.perf_counter()
        [t.start() for t in ths]
        [t.join() for t in ths]
        dt = min(dt, time.perf_counter() - t0)
    print(f"replicas={n}: {n * steps} batches of 8 x 10 s in {dt * 1e3:.1f} ms -> {n * steps * 80 / dt:.0f} audio-s/s "
          f"({dt * 1e3 / (n * steps):.2f} ms per batch)", flush=True)
    print(json.dumps({"batches_in_flight": n, "value": n * steps * 80 / dt, "unit": "audio-s/s",
                      "ms_per_batch": dt * 1e3 / (n * steps), "steps_per_replica": steps}), flush=True)

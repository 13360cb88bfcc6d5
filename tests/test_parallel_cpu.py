"""Host-side data-parallel logic on CPU: LPT partition, length bucketing, and the world-size-2 gather over gloo."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from smalltts_b200 import parallel


def test_partition_lpt_balances_and_covers():
    frames = [75, 15, 60, 22, 75, 30, 45, 15, 8, 70]
    costs = [parallel.utterance_cost(f) for f in frames]
    for world in (1, 2, 4, 8):
        shards = parallel.partition_lpt(costs, world)
        assert sorted(i for s in shards for i in s) == list(range(len(frames)))
        loads = [sum(costs[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(costs)  # LPT bound
    assert parallel.partition_lpt(costs, 2) == parallel.partition_lpt(costs, 2)  # deterministic


def test_length_buckets_bound_padding():
    frames = [75, 74, 70, 40, 38, 15, 15, 14]
    order = sorted(range(len(frames)), key=lambda i: -frames[i])
    mbs = parallel.length_buckets(order, frames, max_batch=4, max_pad_frac=0.2)
    assert sorted(i for m in mbs for i in m) == list(range(len(frames)))
    for m in mbs:
        tmax = max(frames[i] for i in m)
        assert len(m) <= 4 and 1 - sum(frames[i] for i in m) / (tmax * len(m)) <= 0.2 + 1e-9


def _fake_synth(frames):
    def fn(idx):  # deterministic waveform per utterance id
        return [np.full((1, frames[i] * 3200), float(i), np.float32) + np.arange(frames[i] * 3200, dtype=np.float32) * 1e-6
                for i in idx]
    return fn


def _worker(rank, world, port, frames, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = parallel.synthesize_sharded(_fake_synth(frames), frames, rank, world)
    if rank == 0:
        q.put([a[0, :3].tolist() + [a.shape[1]] for a in out])
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_gather_world2_gloo():
    frames = [5, 2, 7, 3, 1]
    want = _fake_synth(frames)(list(range(len(frames))))
    single = parallel.synthesize_sharded(_fake_synth(frames), frames, 0, 1)
    assert all(np.array_equal(a, b) for a, b in zip(single, want))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for g, w in zip(got, want):
        assert g[-1] == w.shape[1] and np.allclose(g[:3], w[0, :3])

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


def test_single_process_multi_device_split_orders_results_and_uses_one_thread_per_worker():
    """SmallTTS(devices=[...]) host logic (parallel.synthesize_on_workers) with fake per-GPU workers."""
    import threading
    import time

    import numpy as np

    from smalltts_b200.parallel import HOP_SIZE, synthesize_on_workers

    frames = [75, 15, 40, 75, 22, 60, 15, 33, 75, 50]
    seen = [set() for _ in range(3)]
    served = [[] for _ in range(3)]

    def make(k):
        def fn(idx):
            seen[k].add(threading.get_ident())
            served[k] += idx
            time.sleep(0.01)
            return [np.full((1, frames[i] * HOP_SIZE), float(i), np.float32) for i in idx]

        return fn

    out = synthesize_on_workers([make(k) for k in range(3)], frames)
    assert [int(a[0, 0]) for a in out] == list(range(len(frames)))
    assert all(a.shape == (1, f * HOP_SIZE) for a, f in zip(out, frames))
    assert sorted(sum(served, [])) == list(range(len(frames)))  # every utterance exactly once
    assert all(len(s) == 1 for s in seen) and len(set.union(*seen)) == 3  # one host thread per worker
    loads = [sum(frames[i] for i in s) for s in served]
    assert max(loads) - min(loads) <= 75  # LPT balance

    # a failing worker surfaces after the others finished; a wrong shape is rejected
    def boom(idx):
        raise RuntimeError("device 1 failed")

    try:
        synthesize_on_workers([make(0), boom], frames)
        raise AssertionError("expected RuntimeError")
    except RuntimeError as e:
        assert "device 1" in str(e)
    try:
        synthesize_on_workers([lambda idx: [np.zeros((1, 5), np.float32) for _ in idx]], [3])
        raise AssertionError("expected ValueError")
    except ValueError:
        pass
    # fewer utterances than workers: idle workers are never called
    calls = []
    out = synthesize_on_workers([lambda idx: calls.append(0) or [np.zeros((1, 2 * HOP_SIZE), np.float32)],
                                 lambda idx: calls.append(1) or []], [2])
    assert calls == [0] and out[0].shape == (1, 2 * HOP_SIZE)

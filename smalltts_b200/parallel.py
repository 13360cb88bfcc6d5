"""Data-parallel batch split across the GPUs of one box.

Utterances are independent, so there is no data-path collective: rank r synthesises its own shard on its own
engine (weights replicated), and only the finished waveforms travel (gathered to rank 0 over the process group:
NCCL on GPUs, gloo in the CPU tests).  The reference has no multi-GPU inference at all (its "batch" is a loop,
infer/onnx.py:143-156, scripts/infer/batch.py:31-45); this module is the host-side scheduler that replaces it.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np

HOP_SIZE = 3200


def utterance_cost(frames: int) -> float:
    """Relative cost model: the vocoder dominates and is linear in frames; a constant covers the condition
    encoder and launch overheads (measured split in DESIGN.md)."""
    return 1.0 * frames + 12.0


def partition_lpt(costs: Sequence[float], n_ranks: int) -> List[List[int]]:
    """Greedy longest-processing-time assignment of items to ranks; deterministic; returns indices per rank,
    each list ordered by decreasing cost (so that consecutive items have similar lengths -> little padding)."""
    if n_ranks < 1:
        raise ValueError("n_ranks must be >= 1")
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0.0] * n_ranks
    out: List[List[int]] = [[] for _ in range(n_ranks)]
    for i in order:
        r = min(range(n_ranks), key=lambda k: (loads[k], k))
        out[r].append(i)
        loads[r] += costs[i]
    return out


# Cost of one engine pass over a padded micro-batch, in frame units: the DiT loop is latency-bound (about the same
# time for 100 or 600 rows) while the vocoder is linear in padded frames.  Measured on B200 (profiles/r02_summary.md):
# ~3.5 ms + 8.5 us per padded frame, i.e. the fixed part is worth ~400 frames.
PASS_FIXED_FRAMES = 400.0


def pass_cost(indices: Sequence[int], frames: Sequence[int]) -> float:
    return PASS_FIXED_FRAMES + len(indices) * max(frames[i] for i in indices)


def shard_cost(indices: Sequence[int], frames: Sequence[int], max_batch: int = 16) -> float:
    return sum(pass_cost(mb, frames) for mb in length_buckets(indices, frames, max_batch=max_batch)) if indices else 0.0


class _ShardScan:
    """Incremental :func:`shard_cost` of a growing contiguous shard (the micro-batching of length_buckets is a left-to-
    right greedy scan, so extending a shard by one utterance is O(1))."""

    def __init__(self, frames: Sequence[int], max_batch: int, max_pad_frac: float = 0.25) -> None:
        self.frames, self.max_batch, self.max_pad_frac = frames, max_batch, max_pad_frac
        self.cost, self.n_cur, self.used, self.tmax = 0.0, 0, 0, 0

    def extended(self, i: int):
        """-> (cost, n_cur, used, tmax) after appending utterance i; does not modify the scan."""
        f = self.frames[i]
        if self.n_cur and not (self.n_cur >= self.max_batch
                               or 1.0 - (self.used + f) / (self.tmax * (self.n_cur + 1)) > self.max_pad_frac):
            return self.cost + self.tmax, self.n_cur + 1, self.used + f, self.tmax
        return self.cost + PASS_FIXED_FRAMES + f, 1, f, f

    def push(self, state) -> None:
        self.cost, self.n_cur, self.used, self.tmax = state


def partition_sorted(frames: Sequence[int], n_ranks: int, max_batch: int = 16) -> List[List[int]]:
    """Contiguous split of the utterances, sorted by decreasing length, into ``n_ranks`` shards that minimises the most
    expensive shard (cost = the engine passes the shard needs, :func:`shard_cost`; bisection on that bound with a greedy
    feasibility scan, O(n log) on the host).  Each rank gets utterances of similar length, so a shard is few, well-
    filled micro-batches: with 64 mixed prompts on 8 GPUs that is ONE pass per GPU, where a length-mixing LPT split
    needs two or three half-empty ones."""
    if n_ranks < 1:
        raise ValueError("n_ranks must be >= 1")
    n = len(frames)
    order = sorted(range(n), key=lambda i: (-frames[i], i))
    if n_ranks == 1 or n == 0:
        return [order] + [[] for _ in range(n_ranks - 1)]

    def split(bound: float):
        """Greedy: fill shard after shard up to `bound`; -> cut points, or None if more than n_ranks shards are needed."""
        cuts, scan = [], _ShardScan(frames, max_batch)
        for pos, i in enumerate(order):
            st = scan.extended(i)
            if st[0] > bound and scan.n_cur:
                cuts.append(pos)
                if len(cuts) >= n_ranks:
                    return None
                scan = _ShardScan(frames, max_batch)
                st = scan.extended(i)
            scan.push(st)
        return cuts

    lo, hi = 0.0, shard_cost(order, frames, max_batch)
    for _ in range(40):
        mid = 0.5 * (lo + hi)
        if split(mid) is None:
            lo = mid
        else:
            hi = mid
        if hi - lo < 0.5:
            break
    cuts = split(hi) or []
    edges = [0] + cuts + [n]
    shards = [order[edges[k] : edges[k + 1]] for k in range(len(edges) - 1)]
    return shards + [[] for _ in range(n_ranks - len(shards))]


def length_buckets(indices: Sequence[int], frames: Sequence[int], max_batch: int = 16,
                   max_pad_frac: float = 0.25) -> List[List[int]]:
    """Split a rank's shard (already sorted by decreasing length) into micro-batches whose padding waste stays
    below ``max_pad_frac`` of the padded size."""
    out: List[List[int]] = []
    cur: List[int] = []
    for i in indices:
        if cur:
            tmax = frames[cur[0]]
            used = sum(frames[j] for j in cur) + frames[i]
            if len(cur) >= max_batch or 1.0 - used / (tmax * (len(cur) + 1)) > max_pad_frac:
                out.append(cur)
                cur = []
        cur.append(i)
    if cur:
        out.append(cur)
    return out


def synthesize_on_workers(workers: Sequence[Callable[[List[int]], List[np.ndarray]]], frames: Sequence[int],
                          max_batch: int = 16) -> List[np.ndarray]:
    """Single-process variant of the batch split (``SmallTTS(devices=[0, 1, ...])``): worker k is bound to GPU k's engine
    (``fn(indices) -> [(1, frames_i*3200)]``) and is driven by its own host thread -- the C-ABI calls release the GIL and
    every engine handle is used by exactly one thread.  Same partitioning as :func:`synthesize_sharded`; the waveforms
    come back in input order.  The first worker exception is re-raised after all threads have finished."""
    from concurrent.futures import ThreadPoolExecutor

    if not workers:
        raise ValueError("no workers")
    shards = partition_sorted(frames, len(workers), max_batch=max_batch)
    results: List[np.ndarray] = [None] * len(frames)  # type: ignore[list-item]

    def run(k: int) -> None:
        for mb in length_buckets(shards[k], frames, max_batch=max_batch):
            for i, a in zip(mb, workers[k](list(mb))):
                if tuple(a.shape) != (1, frames[i] * HOP_SIZE):
                    raise ValueError(f"utterance {i}: expected {(1, frames[i] * HOP_SIZE)}, got {tuple(a.shape)}")
                results[i] = a

    busy = [k for k in range(len(workers)) if shards[k]]
    if len(busy) == 1:
        run(busy[0])
    else:
        with ThreadPoolExecutor(max_workers=len(busy), thread_name_prefix="stts-dev") as pool:
            futs = [pool.submit(run, k) for k in busy]
            errs = [f.exception() for f in futs]
        for e in errs:
            if e is not None:
                raise e
    return results


_PINNED = {}


def _pinned_buffer(n: int):
    """Page-locked fp32 host buffer of at least n elements, cached per process (cudaHostAlloc costs milliseconds)."""
    import torch

    buf = _PINNED.get("buf")
    if buf is None or buf.numel() < n:
        buf = torch.empty(max(n, 1), dtype=torch.float32, pin_memory=True)
        _PINNED["buf"] = buf
    return buf


def synthesize_sharded(synthesize_fn: Callable[[List[int]], List[np.ndarray]], frames: Sequence[int], rank: int,
                       world: int, group=None, gather_to: int = 0, shards: Optional[List[List[int]]] = None,
                       stats: Optional[dict] = None, pinned_out: bool = False):
    """Run ``synthesize_fn`` on this rank's shard and gather every waveform to ``gather_to`` in input order.

    synthesize_fn(indices) -> list of (1, frames_i*3200) float32 arrays (numpy, or torch CUDA tensors when the engine
    leaves its output in HBM) for those utterances (on a GPU rank this is
    ``lambda idx: tts.synthesize_batch([refs[i] for i in idx], ..., device_out=True)``).  Returns the full list on ``gather_to``,
    None elsewhere.  world == 1 needs no process group.  ``shards`` overrides the split (default: :func:`partition_sorted`,
    the same on every rank); ``stats`` receives ``compute_s`` and ``gather_s`` of this rank.  ``pinned_out`` (NCCL only):
    the gathered waveforms are views of ONE page-locked host buffer filled by a single device-to-host copy (about 2.5x
    faster than pageable copies); the buffer is reused by the next call with ``pinned_out``, so copy what must outlive it."""
    import time

    t0 = time.perf_counter()
    if shards is None:
        shards = partition_sorted(frames, world)
    mine = shards[rank]
    results = {}
    for mb in length_buckets(mine, frames):
        for i, a in zip(mb, synthesize_fn(list(mb))):
            if tuple(a.shape) != (1, frames[i] * HOP_SIZE):
                raise ValueError(f"utterance {i}: expected {(1, frames[i] * HOP_SIZE)}, got {tuple(a.shape)}")
            results[i] = a
    on_device = any(type(a).__module__.startswith("torch") and a.is_cuda for a in results.values())
    if on_device:
        import torch

        torch.cuda.synchronize()
    t1 = time.perf_counter()
    if stats is not None:
        stats["compute_s"], stats["gather_s"] = t1 - t0, 0.0

    def to_numpy(a):
        return np.asarray(a.detach().cpu().numpy() if type(a).__module__.startswith("torch") else a, dtype=np.float32)

    if world == 1:
        return [to_numpy(results[i]) for i in range(len(frames))]
    import torch
    import torch.distributed as dist

    # one flat fp32 buffer per rank, sizes are known on every rank from `frames` (no size exchange needed).  When the
    # engine left the waveforms in HBM they are concatenated and sent from there (NCCL over NVLink, no host round trip).
    sizes = [sum(frames[i] for i in s) * HOP_SIZE for s in shards]
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    if on_device and dev.type == "cuda":
        send = torch.cat([results[i].reshape(-1) for i in mine]) if mine else torch.zeros(0, device=dev)
    else:
        flat = np.concatenate([to_numpy(results[i]).ravel() for i in mine]) if mine else np.zeros(0, np.float32)
        send = torch.from_numpy(flat).to(dev)
    if rank == gather_to:
        bufs = [torch.empty(n, dtype=torch.float32, device=dev) if r != rank else send for r, n in enumerate(sizes)]
        reqs = [dist.irecv(bufs[r], src=r, group=group) for r in range(world) if r != rank and sizes[r] > 0]
        for q in reqs:
            q.wait()
        out: List[np.ndarray] = [None] * len(frames)  # type: ignore[list-item]
        if pinned_out and dev.type == "cuda":
            total = sum(sizes)
            host_all = _pinned_buffer(total)
            host_all[:total].copy_(torch.cat([b.reshape(-1) for b in bufs]), non_blocking=True)
            torch.cuda.current_stream().synchronize()
            hosts, base = [], 0
            for r in range(world):
                hosts.append(host_all[base : base + sizes[r]].numpy())
                base += sizes[r]
        else:
            hosts = [bufs[r].cpu().numpy() for r in range(world)]
        for r in range(world):
            off, host = 0, hosts[r]
            for i in shards[r]:
                n = frames[i] * HOP_SIZE
                out[i] = host[off : off + n].reshape(1, n)
                off += n
        if stats is not None:
            stats["gather_s"] = time.perf_counter() - t1
        return out
    if sizes[rank] > 0:
        dist.send(send, dst=gather_to, group=group)
    if stats is not None:
        stats["gather_s"] = time.perf_counter() - t1
    return None

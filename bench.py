#!/usr/bin/env python
"""Benchmark of the smalltts synthesize hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of synthetic input: BASELINE.json configs[1],
batch = 8 x 10 s utterances per GPU (T = 75 latent frames, R = 15 reference frames, P = 120 phonemes), condition
encoder + 4-step DMD denoiser loop + VibeVoice decoder, random-init weights of the reference architecture (seeded;
no checkpoints can be fetched offline).  Metric: audio-seconds generated per wall second (= 1/RTF, the reference's
bench.rs:69-71), aggregated over all GPUs.

  value : device-timed (CUDA events on the engine stream), inputs resident in HBM, output left in HBM
  e2e   : the same through the public Python API / C ABI with pinned HOST buffers; H2D of inputs and D2H of the
          waveform are inside the timed region
  --impl reference : the reference's own CPU path (PyTorch fp32 restatement in oracle/, all host threads) on the SAME
          config: every step runs all 8 utterances x 10 s, one after the other like the reference's batch loop
          (infer/onnx.py:143-156, bench.rs:29); --steps / --warmup are honoured exactly

N > 1 additionally runs BASELINE configs[3] (64 mixed 2-10 s prompts sharded over the N GPUs, smalltts_b200/parallel.py)
after the headline measurement and reports it under "config4" (a STRONG-scaling figure: the job is fixed).

N > 1: one process per GPU (torchrun), utterance batches are independent -> weak scaling, no data-path collective;
NCCL only for the barrier and the max-over-ranks of the timings.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH, T, R, P, STEPS_DMD = 8, 75, 15, 120, 4
AUDIO_S_PER_UTT = T * 3200 / 24000.0
METRIC = "audio_seconds_per_second"
UNIT = "audio-s/s"
WORKLOAD = "configs[1]: batch=8 x 10 s utterances per GPU (T=75, R=15, P=120), DMD 4-step DiT + vocoder"


def tail_algorithmic_bytes(batch: int, frames: int) -> float:
    """fp32 activations, every conv / ConvNeXt layer reads its input once and writes its output
    once.  Vocoder HBM-bound tail = stages with C <= 128 (up3, up4, up5) + head."""
    total = 0.0
    rows_in, c_in = frames * 200, 256  # output of up2
    for r, c in ((4, 128), (2, 64), (2, 32)):
        rows = rows_in * r
        total += rows_in * c_in * 4 + rows * c * 4  # transposed conv: read input, write output
        total += 3 * 2 * rows * c * 4  # three ConvNeXt layers
        rows_in, c_in = rows, c
    total += rows_in * 32 * 4 + rows_in * 4  # head conv 32 -> 1
    return total * batch


def measured_peaks():
    """-> (hbm GB/s, bf16 TFLOP/s sustained, bf16 TFLOP/s burst, source)."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return (float(p["hbm_gbs"]), float(p.get("bf16_tflops_sustained", 1400.0)), float(p.get("bf16_tflops", 1650.0)),
                "measured (MEASURED_PEAKS.json)")
    except Exception:
        return 6650.0, 1400.0, 1650.0, "fallback (B200_PROFILING.md: 6.65 TB/s, ~1.4 PFLOP/s sustained)"


TRAFFIC_FILE = "r02_tail_traffic.json"


def tail_dram_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the tail group's kernels, per launch group, from the ncu capture
    committed under profiles/ (tools/ncu_tail.sh writes it together with the commit it was taken at; same shapes as
    this bench).  -> (bytes, provenance) or (None, None)."""
    try:
        with open(os.path.join(ROOT, "profiles", TRAFFIC_FILE)) as f:
            j = json.load(f)
        return float(j["traffic_bytes"]), f"profiles/{TRAFFIC_FILE} (ncu --set full, captured at commit {j.get('commit', '?')})"
    except Exception:
        return None, None


PRECISION_NOTE = {"fast": "bf16 GEMM / attention operands, fp32 accumulation (stated tolerance 2e-2 rel-L2 vs the fp32 reference)",
                  "tight": "fp16 operands (11-bit significand like TF32), fp32 accumulation (stated tolerance 2.5e-3 rel-L2, measured 7e-4 on waveforms)"}

# algorithmic FLOPs of one denoiser evaluation at B=8, T=75, R=15, P=120
DENOISE_GFLOP_PER_STEP = 177.3


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ------------------------------------------------------------------------------------------ reference arm (CPU)
def cpu_reference_run(steps: int, warmup: int, utterances: int = BATCH, batched: bool = False):
    """The reference's own CPU implementation of the path, restated in oracle/ (PyTorch fp32, torch threads = all
    host cores).  One step = `utterances` utterances of the workload (10 s, R=15, P=120), processed one after the other
    like the reference's batch loop (infer/onnx.py:143-156, bench.rs:29,57-63), or as one padded batch (`batched`: what
    the reference's mask-aware PyTorch model could do but none of its entry points does).
    -> (audio-s/s, seconds per step, threads)."""
    import torch

    from oracle import smalltts_oracle as O
    from smalltts_b200 import synthetic

    torch.set_num_threads(os.cpu_count() or 1)
    sd, vsd = synthetic.dit_state_dict(0), synthetic.vocoder_state_dict(1)
    refs, ids, frames, noise = synthetic.synthetic_inputs(utterances, T, R, P)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        if batched:
            O.synthesize_batch(sd, vsd, refs, ids, frames, noise)
        else:
            for u in range(utterances):
                O.synthesize_batch(sd, vsd, refs[u : u + 1], ids[u : u + 1], frames[u : u + 1], noise[:, u : u + 1])
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    per_step = float(np.mean(times))
    return utterances * AUDIO_S_PER_UTT / per_step, per_step, torch.get_num_threads()


def main_reference(args, rank, world):
    if rank != 0:
        return
    steps, warmup = args.steps, args.warmup
    value, per_step, cores = cpu_reference_run(steps, warmup)
    sample = (f"{steps} steps x {BATCH} utterances x 10 s (T=75,R=15,P=120) after {warmup} warm-up steps; sequential per "
              "utterance like the reference's batch loop")
    batched, _, _ = cpu_reference_run(1, 0, batched=True)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "per_gpu_batch": BATCH, "dmd_steps": STEPS_DMD,
                   "reference_path": "PyTorch fp32 CPU restatement (oracle/); onnxruntime and the .onnx assets are not "
                   "available offline"},
        "rtf": 1.0 / value,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "batched_value": batched,
                         "batched_sample": f"1 step, the {BATCH} utterances as ONE padded batch (not a reference entry point)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------ BASELINE configs[3]
def run_config4(eng, rank, world, local_rank, dist, reps: int = 3):
    """64 mixed 2-10 s prompts (T ~ U{15..75}, R ~ U{8..64}, P = round(1.53 T)) sharded over the ranks
    (smalltts_b200/parallel.py: length-sorted contiguous split, length-bucketed micro-batches per rank), waveforms
    gathered to rank 0 from HBM over NCCL.  The job is fixed, so across N this is STRONG scaling.  Wall clock around the
    whole job including the host-side partition, padding, H2D and the gather; best of `reps` after one warm-up pass
    (which builds the per-shape plans / CUDA graphs)."""
    import torch

    from smalltts_b200 import parallel, synthetic
    from smalltts_b200.infer import SmallTTS

    rng = np.random.default_rng(20260217)
    n = 64
    frames = rng.integers(15, 76, n).tolist()
    refs_n = rng.integers(8, 65, n).tolist()
    phon_n = [int(round(1.53 * f)) for f in frames]
    refs, ids, _, _ = synthetic.synthetic_inputs(n, frames, refs_n, phon_n, steps=1)
    durs = [f * 3200 / 24000 + 1e-3 for f in frames]
    tts = SmallTTS(engine=eng)

    def fn(idx):
        return tts.synthesize_batch([refs[i] for i in idx], [ids[i] for i in idx], [durs[i] for i in idx], seed=7,
                                    device_out=world > 1)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    best, best_stats, out = None, None, None
    for rep in range(1 + reps):
        stats = {}
        barrier()
        t0 = time.perf_counter()
        out = parallel.synthesize_sharded(fn, frames, rank, world, stats=stats, pinned_out=True)
        barrier()
        dt = time.perf_counter() - t0
        if rep > 0 and (best is None or dt < best):
            best, best_stats = dt, stats
    t = torch.tensor([best, best_stats["compute_s"], best_stats["gather_s"] if rank == 0 else 0.0], dtype=torch.float64,
                     device=f"cuda:{local_rank}")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank != 0:
        return None
    assert all(a.shape == (1, f * 3200) and np.isfinite(a).all() for a, f in zip(out, frames))
    shards = parallel.partition_sorted(frames, world)
    pads = [(max(frames[i] for i in mb) * len(mb), sum(frames[i] for i in mb))
            for s in shards for mb in parallel.length_buckets(s, frames)]
    audio_s = sum(frames) * 3200 / 24000
    return {"workload": "configs[3]: 64 mixed 2-10 s prompts sharded over the GPUs (strong scaling: fixed job)",
            "audio_s_per_s": audio_s / t[0].item(), "wall_ms": t[0].item() * 1e3, "audio_seconds": audio_s,
            "compute_ms_max_rank": t[1].item() * 1e3, "gather_ms": t[2].item() * 1e3, "micro_batches": len(pads),
            "micro_batches_per_rank": [len(parallel.length_buckets(s, frames)) for s in shards],
            "padding_frac": 1 - sum(u for _, u in pads) / sum(p for p, _ in pads),
            "timing": f"wall clock incl. host partition/padding, H2D and the gather to rank 0; best of {reps} after a warm-up pass"}


# ------------------------------------------------------------------------------------------ our arm (B200)
def main_ours(args, rank, local_rank, world):
    import torch

    from smalltts_b200 import synthetic
    from smalltts_b200.engine import Engine, pad_batch

    dist = None
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    dev = f"cuda:{local_rank}"
    eng = Engine(local_rank, precision=args.precision)
    eng.load_state_dicts(synthetic.dit_state_dict(0), synthetic.vocoder_state_dict(1))

    # per-rank batch: same shapes, rank-dependent content
    refs, ids, frames, _ = synthetic.synthetic_inputs(BATCH, T, R, P, seed=20260217 + rank)
    ref, ref_len, idt, ph_len = pad_batch(refs, ids, frames)
    d_ref = torch.from_numpy(ref).to(dev)
    d_ids = torch.from_numpy(idt).to(dev)
    d_out = torch.empty(BATCH, T * 3200, dtype=torch.float32, device=dev)
    h_ref, h_ids = eng.pinned(ref.shape), eng.pinned(idt.shape, np.int64)
    h_ref[...] = ref
    h_ids[...] = idt
    h_out = eng.pinned((BATCH, T * 3200))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device(i):
        eng.synthesize(d_ref, ref_len, d_ids, ph_len, frames, T, seed=1000 + i, steps=STEPS_DMD, out=d_out)

    def step_host(i):
        eng.synthesize(h_ref, ref_len, h_ids, ph_len, frames, T, seed=1000 + i, steps=STEPS_DMD, out=h_out)

    def timed(fn, steps, warmup, sample_clocks):
        for i in range(warmup):
            fn(i)
        sampler = ClockSampler(local_rank) if sample_clocks else None
        barrier()
        if sampler:
            sampler.start()
        l0 = Engine.launch_count()
        tail = front = 0.0
        stage = {"cond_enc_ms": 0.0, "denoise_ms": 0.0, "codec_dec_ms": 0.0}
        eng.timer_start()
        w0 = time.perf_counter()
        for i in range(steps):
            fn(warmup + i)
            v = eng.vocoder_ms()
            tail += v["tail_hbm"]
            front += v["front_tensor"]
            tm = eng.timings()
            for k in stage:
                stage[k] += tm[k]
        dev_ms = eng.timer_stop()
        barrier()
        wall_ms = (time.perf_counter() - w0) * 1e3
        launches = Engine.launch_count() - l0
        clocks = sampler.stop() if sampler else None
        t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t[0].item(), t[1].item(), launches, tail / steps, front / steps, {k: v / steps for k, v in stage.items()}, clocks

    dev_ms, wall_ms, launches, tail_ms, front_ms, stage, clocks = timed(step_device, args.steps, args.warmup, True)
    e2e_ms, e2e_wall, _, _, _, _, _ = timed(step_host, args.steps, max(1, args.warmup // 2), False)

    # correctness guard inside the bench: finite, non-trivial output
    a = d_out[:, ::997].float().cpu().numpy()
    assert np.isfinite(a).all() and a.std() > 0, "engine produced non-finite or constant audio"

    audio_s = BATCH * AUDIO_S_PER_UTT * world * args.steps
    value = audio_s / (dev_ms / 1e3)
    e2e_value = audio_s / (e2e_wall / 1e3)
    peak, peak_tf, peak_tf_burst, peak_src = measured_peaks()
    traffic, traffic_src = tail_dram_traffic()
    tail_bytes = tail_algorithmic_bytes(BATCH, T)
    achieved = tail_bytes / (tail_ms / 1e3) / 1e9
    dit_tflops = DENOISE_GFLOP_PER_STEP * STEPS_DMD / stage["denoise_ms"]  # GFLOP / ms = TFLOP/s

    # Informational, not the headline (BASELINE's metric is one batch at a time): throughput with several batches in
    # flight on this GPU -- N engine replicas, one host thread and stream each (tools/bench_concurrent.py).  Runs in a
    # child process after the measurements above, so nothing it does can disturb them; any failure just drops the key.
    in_flight = None
    if rank == 0 and world == 1 and args.in_flight > 1:
        try:
            import subprocess

            tool = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools", "bench_concurrent.py")
            r = subprocess.run([sys.executable, tool, str(max(5, min(args.steps, 20))), str(args.in_flight), str(local_rank)],
                               stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=150)
            rows = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
            if rows:
                in_flight = {"by_batches_in_flight": {str(x["batches_in_flight"]): round(x["value"], 1) for x in rows},
                             "unit": UNIT, "timing": "wall clock around all replica threads (best of two passes), device-resident inputs",
                             "note": "N engine replicas on one GPU; the headline value above is 1 batch at a time"}
        except Exception:  # noqa: BLE001 - strictly optional
            in_flight = None

    # The other precision mode on the same workload (device-resident inputs, CUDA events on the engine stream): the
    # headline build feeds bf16 operands to the tensor cores, the parity build fp16 (11-bit significand like TF32).
    other = None
    if rank == 0 and world == 1 and not args.no_other_precision:
        try:
            oname = "tight" if args.precision == "fast" else "fast"
            eng2 = Engine(local_rank, precision=oname)
            eng2.load_state_dicts(synthetic.dit_state_dict(0), synthetic.vocoder_state_dict(1))
            for i in range(3):
                eng2.synthesize(d_ref, ref_len, d_ids, ph_len, frames, T, seed=1 + i, steps=STEPS_DMD, out=d_out)
            n2 = 10
            eng2.timer_start()
            for i in range(n2):
                eng2.synthesize(d_ref, ref_len, d_ids, ph_len, frames, T, seed=10 + i, steps=STEPS_DMD, out=d_out)
            ms2 = eng2.timer_stop()
            other = {"precision": oname, "value": BATCH * AUDIO_S_PER_UTT * n2 / (ms2 / 1e3), "unit": UNIT,
                     "ms_per_step": ms2 / n2, "steps": n2, "warmup": 3}
            eng2.close()
        except Exception as exc:  # noqa: BLE001 - informational
            other = {"error": repr(exc)}

    config4 = None
    if not args.no_config4:
        try:
            config4 = run_config4(eng, rank, world, local_rank, dist)
        except Exception as exc:  # noqa: BLE001 - the headline line must still be printed
            config4 = {"error": repr(exc)} if rank == 0 else None

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, per_step, cores = cpu_reference_run(2, 1)
        vb, _, _ = cpu_reference_run(1, 0, batched=True)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"2 steps x {BATCH} utterances x 10 s (T=75,R=15,P=120) after 1 warm-up step; reference "
                         "semantics (sequential per utterance), PyTorch fp32 oracle",
               "batched_value": vb,
               "batched_sample": f"1 step, the {BATCH} utterances as ONE padded batch (not a reference entry point)"}
    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "fast" else "fp16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "per_gpu_batch": BATCH, "dmd_steps": STEPS_DMD, "noise": "on-device Philox",
                       "precision": f"{args.precision}: {PRECISION_NOTE[args.precision]}",
                       "l2": "no explicit flush: per-step working set (1.3 GB bf16 weights + >1 GB activations) >> 126 MB L2",
                       "weights": "seeded random init of the reference architecture (328 M DiT + 344 M vocoder params)"},
            "rtf": (dev_ms / 1e3) / audio_s,
            "wall_ms_per_step": wall_ms / args.steps,
            "stage_ms": stage,
            "gpu_launches": launches,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_wall / args.steps,
                    "h2d_bytes_per_step": int(ref.nbytes + idt.nbytes), "d2h_bytes_per_step": int(BATCH * T * 3200 * 4),
                    "api": "Engine.synthesize / stts_synthesize with pinned host buffers, wall clock around the call"},
            "roofline": {"bound": "hbm", "kernel": "vocoder tail (stages up3..up5 with C<=128 + head): token-mixer + "
                         "tcgen05 FFN GEMMs + transposed-conv GEMMs, CUDA events around the group on the engine stream",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "algorithmic_bytes": tail_bytes, "ms": tail_ms,
                         "traffic": traffic, "traffic_source": traffic_src, "front_ms": front_ms},
            # second regime of the path: the 4 denoiser evaluations are tensor-core work
            "roofline_tensor": {"bound": "tensor", "kernel": "DMD loop: 4 denoiser evaluations (tcgen05 GEMMs + attention), "
                                "CUDA events around the loop on the engine stream",
                                "achieved": dit_tflops, "peak": peak_tf, "unit": "TFLOP/s", "frac": dit_tflops / peak_tf,
                                "peak_burst": peak_tf_burst, "frac_of_burst": dit_tflops / peak_tf_burst,
                                "algorithmic_gflop": DENOISE_GFLOP_PER_STEP * STEPS_DMD, "ms": stage["denoise_ms"]},
            "cpu_baseline": cpu,
        }
        if other is not None:
            out["other_precision"] = other
        if in_flight is not None:
            out["in_flight"] = in_flight
        if config4 is not None:
            out["config4"] = config4
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="fast", choices=["fast", "tight"],
                    help="fast: bf16 tensor-core operands (headline); tight: the fp16-operand parity build")
    ap.add_argument("--no-other-precision", action="store_true", help="skip the short run of the other precision mode")
    ap.add_argument("--no-config4", action="store_true", help="skip the configs[3] (64 mixed prompts, sharded) figure")
    ap.add_argument("--in-flight", type=int, default=2,
                    help="also report throughput with up to this many batches in flight on one GPU (N=1 only; 0 = skip)")
    args = ap.parse_args()
    rank, local_rank, world = dist_env()
    if args.impl == "reference":
        main_reference(args, rank, world)
    else:
        if args.warmup < 3:
            args.warmup = 3
        main_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()

"""Batch synthesis on the B200 engine: the reference's scripts/infer/batch.py, but the prompts run as ONE ragged
engine call (and are sharded over the GPUs of the box when launched under torchrun) instead of a Python loop.

    python scripts/batch.py --items items.json [--outdir out] [model args of scripts/clone.py]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/batch.py --items items.json

items.json: [{"filename": "a.wav", "text": "...", "tokens": [..optional ids..], "duration": optional seconds}, ...]
(the reference reads assets/test_audio/transcriptions.json and a fixed list of four texts, batch.py:15-24).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from clone import add_model_args, load_tts, tokens_for  # noqa: E402

from smalltts_b200.infer import estimate_duration  # noqa: E402
from smalltts_b200.serve import decode_wav, encode_wav  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--items", required=True)
    ap.add_argument("--outdir", default="out")
    add_model_args(ap)
    args = ap.parse_args()
    with open(args.items) as f:
        items = json.load(f)
    base = os.path.dirname(os.path.abspath(args.items))
    os.makedirs(args.outdir, exist_ok=True)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    tts = load_tts(args)

    refs, toks, durs = [], [], []
    for it in items:
        with open(os.path.join(base, it["filename"]), "rb") as fh:
            y, sr = decode_wav(fh.read())
        refs.append(tts.clone_voice(y, sample_rate=sr))
        tok = it.get("tokens") or tokens_for(it.get("text", ""), "")
        toks.append(list(map(int, tok)))
        durs.append(float(it.get("duration") or estimate_duration(it.get("text") or " " * len(tok))))

    # utterances are independent: each rank synthesises its shard (length-bucketed micro-batches) on its own engine and
    # rank 0 gathers the waveforms straight from HBM (smalltts_b200/parallel.py; no data-path collective)
    from smalltts_b200.infer import frames_for
    from smalltts_b200.parallel import synthesize_sharded

    audio = synthesize_sharded(
        lambda idx: tts.synthesize_batch([refs[i] for i in idx], [toks[i] for i in idx], [durs[i] for i in idx],
                                         device_out=world > 1),
        [frames_for(d) for d in durs], rank, world)
    if rank == 0:
        for it, a in zip(items, audio):
            out_path = os.path.join(args.outdir, os.path.splitext(os.path.basename(it["filename"]))[0] + "_gen.wav")
            with open(out_path, "wb") as fh:
                fh.write(encode_wav(a.squeeze(), 24_000))
            print("  ->", out_path)

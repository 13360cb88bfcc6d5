#!/usr/bin/env python
"""Convert the reference's weight files into this package's flat `.sttsw` containers (mmap-able, checked against the
architecture at conversion time, optionally bf16 for the matrices -> half the size).

    python tools/convert_weights.py dit     assets/dmd/condition_encoder.onnx assets/dmd/denoiser.onnx -o dit.sttsw
    python tools/convert_weights.py decoder assets/codec/decoder.onnx -o decoder.sttsw [--bf16]
    python tools/convert_weights.py encoder assets/codec/encoder.onnx -o encoder.sttsw
    python tools/convert_weights.py dit     ckpt/student.pt -o dit.sttsw          # trainer checkpoints work too

Then: SmallTTS("dit.sttsw", None, "decoder.sttsw", codec_encoder_path="encoder.sttsw")."""
from __future__ import annotations

import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main(argv=None) -> int:
    from smalltts_b200 import synthetic, weights

    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("model", choices=["dit", "decoder", "encoder"])
    ap.add_argument("files", nargs="+", help=".onnx / .pt / .safetensors / .sttsw (several files are merged)")
    ap.add_argument("-o", "--out", required=True)
    ap.add_argument("--bf16", action="store_true", help="store >= 2-D tensors as bfloat16 (the engine rounds them anyway)")
    args = ap.parse_args(argv)
    specs, rank = {"dit": (synthetic.dit_specs(), weights.dit_exec_rank), "decoder": (synthetic.vocoder_specs(), None),
                   "encoder": (synthetic.encoder_specs(), None)}[args.model]
    sd = weights.load_model_weights(args.files, specs, args.model, exec_rank=rank)
    weights.save_packed(args.out, sd, dtype="bfloat16" if args.bf16 else "float32")
    n = sum(int(v.size if hasattr(v, "size") and not callable(v.size) else v.numel()) for v in sd.values())
    print(f"{args.out}: {len(sd)} tensors, {n:,} parameters, {os.path.getsize(args.out) / 1e6:.1f} MB")
    return 0


if __name__ == "__main__":
    sys.exit(main())

"""Voice cloning on the B200 engine: the reference's scripts/infer/clone.py with the imports re-pointed.

    python scripts/clone.py --wav ref.wav --text "hello" [--duration 3.0] [--out out/clone.wav]
        [--dit assets/dmd/condition_encoder.onnx --denoiser assets/dmd/denoiser.onnx
         --decoder assets/codec/decoder.onnx --encoder assets/codec/encoder.onnx]
        [--tokens 12,7,33]      # instead of --text when the espeak phonemizer is not installed
        [--synthetic]           # seeded random weights (no checkpoint available offline)

Differences from the reference script: WAV I/O through the stdlib (16-bit PCM / float32; ``soundfile`` is optional),
resampling and the codec encoder run on the GPU, and there is no asset download (no network).
"""
from __future__ import annotations

import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from smalltts_b200.infer import SmallTTS, estimate_duration  # noqa: E402
from smalltts_b200.serve import decode_wav, encode_wav  # noqa: E402


def tokens_for(text, tokens):
    if tokens:
        return [int(t) for t in tokens.split(",")]
    from smalltts.data.phonemization.phonemes import get_token_ids  # the reference's espeak front-end

    return get_token_ids(text)


def load_tts(args) -> SmallTTS:
    if args.synthetic:
        return SmallTTS.synthetic(encoder_seed=2)
    if all(str(p).startswith("assets/") for p in (args.dit, args.denoiser, args.decoder, args.encoder)):
        from smalltts_b200.assets import ensure_assets  # like the reference scripts (clone.py:14)

        ensure_assets(["codec", "dmd"])
    return SmallTTS(args.dit, args.denoiser, args.decoder, codec_encoder_path=args.encoder)


def add_model_args(ap: argparse.ArgumentParser) -> None:
    ap.add_argument("--dit", default="assets/dmd/condition_encoder.onnx")
    ap.add_argument("--denoiser", default="assets/dmd/denoiser.onnx")
    ap.add_argument("--decoder", default="assets/codec/decoder.onnx")
    ap.add_argument("--encoder", default="assets/codec/encoder.onnx")
    ap.add_argument("--synthetic", action="store_true", help="seeded random weights instead of checkpoints")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--wav", required=True, help="reference audio file")
    ap.add_argument("--text", default="", help="text to speak")
    ap.add_argument("--tokens", default="", help="comma-separated phoneme token ids (bypasses the phonemizer)")
    ap.add_argument("--duration", type=float, default=None, help="duration in seconds (auto if omitted)")
    ap.add_argument("--out", default="out/clone.wav")
    add_model_args(ap)
    args = ap.parse_args()

    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    print("loading")
    tts = load_tts(args)
    with open(args.wav, "rb") as fh:
        y, sr = decode_wav(fh.read())
    print("encoding reference audio")
    ref_latents = tts.clone_voice(y, sample_rate=sr)  # mono mix + resample_hq + Encoder.encode, on the GPU
    tokens = tokens_for(args.text, args.tokens)
    duration = args.duration or estimate_duration(args.text or " " * len(tokens))
    print(f"generating ({duration:.1f}s)")
    audio = tts.synthesize(ref_latents, tokens, duration)
    with open(args.out, "wb") as fh:
        fh.write(encode_wav(audio.squeeze(), 24_000))
    print(args.out)

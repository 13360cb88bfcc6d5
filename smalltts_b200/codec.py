"""Drop-in for ``smalltts.codec.onnx`` (codec/onnx.py:34-75) on the B200 engine."""
from __future__ import annotations

from typing import Iterable, Optional

import numpy as np

from .engine import Engine
from .infer import load_state_dict_file


class Decoder:
    """VibeVoice acoustic-tokenizer decoder: latents (B,T,64) -> audio (B,1,T*3200)  (codec/onnx.py:34-53)."""

    def __init__(self, path: str = "assets/codec/decoder.safetensors", providers: Optional[Iterable[str]] = None, *,
                 engine: Optional[Engine] = None) -> None:
        if engine is None:
            raise NotImplementedError(
                "a standalone Decoder needs the DiT weights too in this version; pass engine=SmallTTS(...).engine")
        self.engine = engine

    def decode(self, latents):
        import torch

        is_t = isinstance(latents, torch.Tensor)
        x = latents.detach().cpu().numpy() if is_t and not latents.is_cuda else latents
        y = self.engine.decode(x if is_t and latents.is_cuda else np.asarray(x, dtype=np.float32))
        y = y if isinstance(y, torch.Tensor) else torch.from_numpy(y)
        return y[:, None, :]


class Encoder:
    """VibeVoice acoustic-tokenizer encoder: audio (B,1,N) @ 24 kHz -> latents (B, N // 3200, 64)
    (codec/onnx.py:56-75; the clone path, scripts/infer/clone.py:36).  The engine must have been given the encoder
    weights (``SmallTTS(..., codec_encoder_path=...)`` or ``state_dicts=(dit, decoder, encoder)``)."""

    def __init__(self, path: str = "assets/codec/encoder.safetensors", providers: Optional[Iterable[str]] = None, *,
                 engine: Optional[Engine] = None) -> None:
        if engine is None:
            raise NotImplementedError(
                "a standalone Encoder needs an engine in this version; pass engine=SmallTTS(...).engine")
        self.engine = engine

    def encode(self, audio):
        import torch

        is_t = isinstance(audio, torch.Tensor)
        x = audio.detach().cpu().numpy() if is_t and not audio.is_cuda else audio
        y = self.engine.encode_audio(x if is_t and audio.is_cuda else np.asarray(x, dtype=np.float32))
        return y if isinstance(y, torch.Tensor) else torch.from_numpy(y)

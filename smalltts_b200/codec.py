"""Drop-in for ``smalltts.codec.onnx`` (codec/onnx.py:34-75) on the B200 engine.

``Decoder(path)`` / ``Encoder(path)`` build a codec-only engine from a weight file (``.onnx`` initialisers,
``.safetensors``, ``.pt``, ``.sttsw``; see ``smalltts_b200/weights.py``) exactly like the reference's standalone use
(scripts/train/dmd2/sv.py:24 builds a bare ``Decoder``); ``engine=`` shares the engine of an existing ``SmallTTS``."""
from __future__ import annotations

from typing import Iterable, Optional

import numpy as np

from .engine import Engine


def _to_engine_input(x):
    import torch

    if isinstance(x, torch.Tensor):
        return x if x.is_cuda else x.detach().cpu().numpy()
    return np.asarray(x, dtype=np.float32)


def _as_tensor(y):
    import torch

    return y if isinstance(y, torch.Tensor) else torch.from_numpy(y)


class Decoder:
    """VibeVoice acoustic-tokenizer decoder: latents (B,T,64) -> audio (B,1,T*3200)  (codec/onnx.py:34-53)."""

    def __init__(self, path: str = "assets/codec/decoder.onnx", providers: Optional[Iterable[str]] = None, *,
                 engine: Optional[Engine] = None, state_dict=None, device: int = 0) -> None:
        if engine is None:
            if state_dict is None:
                from . import synthetic
                from .weights import load_model_weights

                state_dict = load_model_weights([path], synthetic.vocoder_specs(), "codec decoder")
            engine = Engine(device)
            engine.load_state_dicts(None, state_dict, None)
        self.engine = engine

    def decode(self, latents):
        return _as_tensor(self.engine.decode(_to_engine_input(latents)))[:, None, :]


class Encoder:
    """VibeVoice acoustic-tokenizer encoder: audio (B,1,N) @ 24 kHz -> latents (B, N // 3200, 64)
    (codec/onnx.py:56-75; the clone path, scripts/infer/clone.py:36)."""

    def __init__(self, path: str = "assets/codec/encoder.onnx", providers: Optional[Iterable[str]] = None, *,
                 engine: Optional[Engine] = None, state_dict=None, device: int = 0) -> None:
        if engine is None:
            if state_dict is None:
                from . import synthetic
                from .weights import load_model_weights

                state_dict = load_model_weights([path], synthetic.encoder_specs(), "codec encoder")
            engine = Engine(device)
            engine.load_state_dicts(None, None, state_dict)
        self.engine = engine

    def encode(self, audio):
        return _as_tensor(self.engine.encode_audio(_to_engine_input(audio)))

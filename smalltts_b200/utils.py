"""Drop-in for ``smalltts.infer.utils`` (infer/utils.py:7-23): the high-quality resampler of the clone path, on the
B200 engine instead of torchaudio (same filter: Kaiser-windowed sinc, lowpass_filter_width 1024, rolloff 0.94)."""
from __future__ import annotations

from typing import Optional

from .engine import Engine

_engine: Optional[Engine] = None


def _default_engine() -> Engine:
    global _engine
    if _engine is None:
        _engine = Engine(0)  # resampling needs no model weights
    return _engine


def resample_hq(x, sr: int, target: int, engine: Optional[Engine] = None):
    """x: torch tensor (..., N) (cpu or cuda) -> (..., ceil(N * target / sr)), same device.  ``sr == target``
    returns ``x`` itself like the reference (infer/utils.py:20-21)."""
    import torch

    if sr == target:
        return x
    eng = engine or _default_engine()
    lead = x.shape[:-1]
    flat = x.reshape(-1, x.shape[-1]).to(torch.float32)
    y = eng.resample(flat if flat.is_cuda else flat.numpy(), sr, target)
    y = y if isinstance(y, torch.Tensor) else torch.from_numpy(y)
    return y.reshape(*lead, y.shape[-1]).to(x.dtype)

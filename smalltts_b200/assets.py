"""Drop-in for ``smalltts.assets.ensure`` (assets/ensure.py:20-41): make sure ``assets/<folder>`` exists, fetching it from
the Hub repo ``smallbraineng/smalltts`` when it does not.  The engine reads the downloaded ``.onnx`` files directly
(``smalltts_b200/weights.py``).  Offline (no ``huggingface_hub`` or no network) the error says which files to provide."""
from __future__ import annotations

import os
from pathlib import Path

REPO = "smallbraineng/smalltts"  # assets/ensure.py:7


def ensure_assets(paths, root: str = "assets") -> None:
    if isinstance(paths, (list, tuple, set)):
        for p in paths:
            ensure_assets(p, root)
        return
    folder = str(paths).strip("/ ")
    if not folder or (Path(root) / folder).exists():
        return
    try:
        from huggingface_hub import snapshot_download

        snapshot_download(repo_id=REPO, allow_patterns=[f"{folder}/*"], local_dir=str(Path(root)),
                          max_workers=os.cpu_count() or 8)
    except Exception as exc:  # no package, no network, no such folder
        raise FileNotFoundError(
            f"{root}/{folder} is missing and could not be fetched from the Hub repo {REPO} ({type(exc).__name__}: {exc}). "
            f"Copy the reference's assets there (dmd/condition_encoder.onnx, dmd/denoiser.onnx, codec/decoder.onnx, "
            f"codec/encoder.onnx, tryme/latents.npy) or pass explicit weight paths to SmallTTS(...).") from exc
    if not (Path(root) / folder).exists():
        raise FileNotFoundError(f"{root}/{folder}: the Hub repo {REPO} has no such folder")

"""Drop-in for the reference's ``smalltts.infer.onnx`` (infer/onnx.py:1-159): same module constants, same
``SmallTTS`` constructor / ``synthesize`` / ``forward`` / ``__call__`` signatures, but every operator runs in the
B200 engine behind the C ABI (no onnxruntime, no PyTorch ops, no CPU fallback).

Differences a caller can see, all additive:
  * the three path arguments point at weight files instead of ``.onnx`` graphs: ``cond_encoder_path`` /
    ``denoiser_path`` name the DiT checkpoint(s) -- ``.pt`` / ``.safetensors`` / ``.sttsw`` with
    ``DiTModel.state_dict()`` keys, or the reference's own ``condition_encoder.onnx`` + ``denoiser.onnx``, whose
    initialisers are read directly (``smalltts_b200/weights.py``; both graphs are exports of that one model) -- and
    ``codec_decoder_path`` the VibeVoice decoder weights (HF ``state_dict`` keys or ``decoder.onnx``);
  * ``synthesize_batch`` runs a ragged batch in ONE engine call (the reference loops, infer/onnx.py:143-156);
  * ``noise=`` / ``seed=`` make the DMD loop reproducible (the reference draws from the global numpy RNG).
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

from .engine import Engine, pad_batch

SAMPLE_RATE = 24_000  # infer/onnx.py:11
HOP_SIZE = 3_200  # infer/onnx.py:12
NUM_STEPS = 4  # infer/onnx.py:13
CHARS_PER_SECOND = 11.5  # infer/onnx.py:14


def estimate_duration(text: str, min_sec: float = 0.5, max_sec: float = 30.0) -> float:
    """infer/onnx.py:17-18."""
    return max(min_sec, min(len(text) / CHARS_PER_SECOND, max_sec))


def frames_for(duration_sec: float) -> int:
    """infer/onnx.py:84."""
    return max(1, int(duration_sec * SAMPLE_RATE / HOP_SIZE))


from .weights import dit_exec_rank, load_model_weights, load_state_dict_file  # noqa: F401  (re-exported)


def _tokens(x) -> List[int]:
    if isinstance(x, str):
        try:  # the reference's own text front-end (espeak), when that package is installed
            from smalltts.data.phonemization.phonemes import get_token_ids
        except Exception as exc:  # pragma: no cover - depends on the host environment
            raise RuntimeError(
                "string inputs need the reference phonemizer (smalltts.data.phonemization); pass token ids instead"
            ) from exc
        return get_token_ids(x)
    return list(map(int, x))


class SmallTTS:
    """DMD 4-step inference on the B200 engine; mirrors infer/onnx.py:50-159."""

    def __init__(
        self,
        cond_encoder_path: str = "assets/dmd/condition_encoder.onnx",  # the reference's defaults (infer/onnx.py:55-57)
        denoiser_path: Optional[str] = "assets/dmd/denoiser.onnx",
        codec_decoder_path: str = "assets/codec/decoder.onnx",
        providers: Optional[Iterable[str]] = None,  # accepted for call compatibility; the engine is CUDA-only
        *,
        codec_encoder_path: Optional[str] = None,  # clone path (codec/onnx.py:56-75): encoder.onnx / HF state_dict file
        device: int = 0,
        devices: Optional[Sequence[int]] = None,  # several GPUs from one process: replicas + batch split
        state_dicts: Optional[tuple] = None,
        num_steps: int = NUM_STEPS,
        seed: Optional[int] = None,
        shape_buckets: Optional[Sequence[int]] = None,  # (R, P, T) multiples the padded batch shape is rounded up to
        engine: Optional[Engine] = None,  # adopt an engine whose weights are already loaded (no second copy)
        precision: str = "fast",  # "tight": the fp16-operand parity build of the same kernels (engine.Engine)
    ) -> None:
        self.num_steps = num_steps
        self.shape_buckets = tuple(int(x) for x in shape_buckets) if shape_buckets is not None else None
        if self.shape_buckets is not None and (len(self.shape_buckets) != 3 or min(self.shape_buckets) < 1):
            raise ValueError("shape_buckets must be three positive integers (R, P, T multiples)")
        # the reference draws fresh, unseeded noise per request (infer/onnx.py:104): without an explicit seed the base of
        # this instance's Philox streams is random, so replicas and restarts do not replay each other
        self._seed = int.from_bytes(os.urandom(7), "little") if seed is None else int(seed)
        self._calls = 0
        devs = [int(d) for d in devices] if devices is not None else [int(device)]
        if not devs or len(set(devs)) != len(devs):
            raise ValueError("devices must be a non-empty list of distinct CUDA ordinals")
        if engine is not None:
            if devices is not None:
                raise ValueError("engine= adopts ONE loaded engine; it cannot be combined with devices=")
            self.engine, self._replicas = engine, []
            return
        if state_dicts is None:
            from . import synthetic

            # both DiT graphs of the reference are exports of one DiTModel (infer/onnx.py:60-62): merge their tensors
            dit_files = [cond_encoder_path] + ([denoiser_path] if denoiser_path is not None else [])
            state_dicts = (load_model_weights(dit_files, synthetic.dit_specs(), "DiTModel", exec_rank=dit_exec_rank),
                           load_model_weights([codec_decoder_path], synthetic.vocoder_specs(), "codec decoder"))
            if codec_encoder_path is not None:
                state_dicts += (load_model_weights([codec_encoder_path], synthetic.encoder_specs(), "codec encoder"),)
        self.engine = Engine(devs[0], precision=precision)
        self.engine.load_state_dicts(*state_dicts)
        # one replica (own engine, own weights copy, own host thread per call) per extra GPU; utterances are independent
        self._replicas = [SmallTTS(state_dicts=state_dicts, device=d, num_steps=num_steps, seed=self._seed + 7919 * (k + 1),
                                   shape_buckets=shape_buckets, precision=precision) for k, d in enumerate(devs[1:])]

    @classmethod
    def synthetic(cls, dit_seed: int = 0, vocoder_seed: int = 1, encoder_seed: Optional[int] = None, **kw) -> "SmallTTS":
        """Engine on seeded random weights of the reference architecture (no checkpoint can be fetched offline)."""
        from . import synthetic

        sds = (synthetic.dit_state_dict(dit_seed), synthetic.vocoder_state_dict(vocoder_seed))
        if encoder_seed is not None:
            sds += (synthetic.encoder_state_dict(encoder_seed),)
        return cls(state_dicts=sds, **kw)

    def clone_voice(self, wav, sample_rate: int = SAMPLE_RATE):
        """scripts/infer/clone.py:27-36: wav (N,), (1,N) or (channels,N) at `sample_rate` -> reference latents (R,64).
        Down-mix to mono (clone.py:29-30), ``resample_hq`` to 24 kHz (infer/utils.py:7-23) and the codec encoder all
        run on the engine; the resampled audio never returns to the host."""
        import torch

        x = torch.as_tensor(np.asarray(wav, dtype=np.float32))
        x = x.reshape(1, -1) if x.ndim == 1 else x.mean(dim=0, keepdim=True)
        x = x.to(f"cuda:{self.engine.device}")
        if sample_rate != SAMPLE_RATE:
            x = self.engine.resample(x, sample_rate, SAMPLE_RATE)
        return self.engine.encode_audio(x)[0].cpu().numpy()

    # ------------------------------------------------------------------ reference API
    def synthesize(self, ref_latents: np.ndarray, phoneme_ids: List[int], duration_sec: float,
                   noise: Optional[np.ndarray] = None) -> np.ndarray:
        """infer/onnx.py:68-129: (R,64) f32, token ids, seconds -> (1, samples) f32 @ 24 kHz."""
        if noise is not None:
            noise = np.asarray(noise, dtype=np.float32)
            if noise.ndim == 3:  # (steps, T, 64) -> (steps, 1, T, 64)
                noise = noise[:, None]
        return self.synthesize_batch([ref_latents], [phoneme_ids], [duration_sec], noise=noise)[0]

    def synthesize_batch(self, ref_latents: Sequence, phoneme_ids: Sequence[Sequence[int]],
                         durations: Sequence[float], noise=None, seed: Optional[int] = None,
                         device_out: bool = False, shape_buckets: Optional[Sequence[int]] = None) -> List[np.ndarray]:
        """Ragged batch in one engine call.  noise: optional (steps, B, Tmax, 64).  Returns [(1, frames_i*3200)]
        (numpy; with ``device_out`` torch CUDA views of the engine's output, for device-side gathers)."""
        if not (len(ref_latents) == len(phoneme_ids) == len(durations)) or len(durations) == 0:
            raise ValueError("ref_latents, phoneme_ids and durations must be equally long and non-empty")
        frames = [frames_for(d) for d in durations]
        if self._replicas and len(frames) > 1 and not device_out:
            return self._synthesize_on_devices(ref_latents, phoneme_ids, durations, frames, noise, seed)
        T = max(frames)
        rb, pb, tb = shape_buckets or self.shape_buckets or (1, 1, 1)
        if noise is None:  # supplied noise fixes T; the on-device stream does not care
            T = -(-T // tb) * tb
        ref, ref_len, ids, ph_len = pad_batch(ref_latents, phoneme_ids, frames, rb, pb)
        if seed is None:
            seed = self._seed + self._calls
        self._calls += 1
        if device_out:
            import torch

            dev = f"cuda:{self.engine.device}"
            audio = self.engine.synthesize(torch.from_numpy(ref).to(dev), ref_len, torch.from_numpy(ids).to(dev), ph_len,
                                           frames, T, noise=None if noise is None else torch.as_tensor(noise).to(dev),
                                           seed=seed, steps=self.num_steps)
            return [audio[i : i + 1, : frames[i] * HOP_SIZE] for i in range(len(frames))]
        audio = self.engine.synthesize(ref, ref_len, ids, ph_len, frames, T, noise=noise, seed=seed,
                                       steps=self.num_steps)
        return [audio[i : i + 1, : frames[i] * HOP_SIZE].copy() for i in range(len(frames))]

    def _synthesize_on_devices(self, ref_latents, phoneme_ids, durations, frames, noise, seed) -> List[np.ndarray]:
        """``devices=[...]``: longest-processing-time split of the utterances over the GPUs, length-bucketed micro-batches
        per GPU, one host thread per GPU (parallel.synthesize_on_workers).  With ``noise`` the result is independent of
        the split; the on-device Philox streams are keyed per (seed, device, micro-batch) and are not."""
        from .parallel import synthesize_on_workers

        noise = None if noise is None else np.asarray(noise, dtype=np.float32)

        def worker(tts: "SmallTTS", k: int):
            calls = [0]

            def run(idx: List[int]) -> List[np.ndarray]:
                tl = max(frames[i] for i in idx)
                nz = None if noise is None else np.ascontiguousarray(noise[:, idx, :tl])
                sd = None if seed is None else int(seed) + 7919 * k + calls[0]
                calls[0] += 1
                return SmallTTS.synthesize_batch(tts, [ref_latents[i] for i in idx], [phoneme_ids[i] for i in idx],
                                                 [durations[i] for i in idx], noise=nz, seed=sd)

            return run

        members = [self] + self._replicas
        saved, self._replicas = self._replicas, []  # the primary serves its own shard locally
        try:
            return synthesize_on_workers([worker(t, k) for k, t in enumerate(members)], frames)
        finally:
            self._replicas = saved

    def synthesize_teacher(self, ref_latents: Sequence, phoneme_ids: Sequence[Sequence[int]], durations: Sequence[float],
                           steps: int = 128, cfg_scale_text: float = 2.0, cfg_scale_speaker: float = 1.5, noise=None,
                           seed: Optional[int] = None) -> List[np.ndarray]:
        """Teacher path of the latency/quality sweep (BASELINE config 5): `steps` DDIM steps with the 3-way CFG of
        scripts/train/dmd2/distill.py:60-134, then the same vocoder.  noise: optional start x_1 (B, Tmax, 64)."""
        if not (len(ref_latents) == len(phoneme_ids) == len(durations)) or len(durations) == 0:
            raise ValueError("ref_latents, phoneme_ids and durations must be equally long and non-empty")
        frames = [frames_for(d) for d in durations]
        T = max(frames)
        ref, ref_len, ids, ph_len = pad_batch(ref_latents, phoneme_ids, frames)
        if seed is None:
            seed = self._seed + self._calls
        self._calls += 1
        cond3 = self.engine.encode_conditions_cfg(ref, ref_len, ids, ph_len)
        try:
            lat = self.engine.sample_teacher(cond3, frames, T, steps=steps, cfg_text=cfg_scale_text,
                                             cfg_speaker=cfg_scale_speaker, noise=noise, seed=seed)
        finally:
            cond3.free()
        audio = self.engine.decode(lat)
        return [audio[i : i + 1, : frames[i] * HOP_SIZE].copy() for i in range(len(frames))]

    def forward(self, conditionings: List, transcriptions: list, texts: list, duration_sec: float = 3.0) -> List:
        """infer/onnx.py:131-157: tokens = transcription tokens + text tokens, one shared duration; returns a list of
        torch tensors (1, samples).  Unlike the reference loop this is a single batched engine call."""
        import torch

        toks = [_tokens(tr) + _tokens(tx) for tr, tx in zip(transcriptions, texts)]
        conds = list(conditionings)[: len(toks)]
        if not toks:
            return []
        out = self.synthesize_batch(conds, toks, [duration_sec] * len(toks))
        return [torch.from_numpy(a) for a in out]

    __call__ = forward

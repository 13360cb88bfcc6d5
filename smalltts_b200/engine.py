"""Host-side handle on the C-ABI engine: one engine per GPU, not thread-safe (like the reference's
``Arc<Mutex<Pipeline>>``, server/src/main.rs:25).  Tensors may be numpy arrays (host memory, copied by the
library) or torch CUDA tensors (device memory, used in place); control metadata is always host."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _cabi

HOP_SIZE = 3200
LATENT_DIM = 64


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _buf(x, dtype, name: str):
    """-> (pointer, mem, keepalive).  numpy -> host pointer; torch cuda tensor -> device pointer."""
    if _is_torch(x):
        import torch

        want = {np.float32: torch.float32, np.int64: torch.int64}[dtype]
        if x.dtype != want:
            raise ValueError(f"{name}: expected dtype {want}, got {x.dtype}")
        x = x.contiguous()
        if x.is_cuda:
            # the engine works on its own non-blocking CUDA stream: whatever torch still has in flight for this tensor
            # (the kernel that produces it, the H2D copy behind .cuda()) must be complete before the engine reads it
            torch.cuda.current_stream(x.device).synchronize()
            return C.c_void_p(x.data_ptr()), _cabi.MEM_DEVICE, x
        x = x.numpy()
    a = np.ascontiguousarray(x, dtype=dtype)
    return C.c_void_p(a.ctypes.data), _cabi.MEM_HOST, a


def _i64(v: Sequence[int]) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(list(v), dtype=np.int64))


class Conditions:
    """Device-resident output of the condition encoder (== the five tensors returned by
    condition_encoder.onnx, infer/onnx.py:94-96).  Reusable across prompts that share a batch of voices."""

    def __init__(self, engine: "Engine", handle, B: int, R: int, P: int):
        self._engine, self._h, self.B, self.R, self.P = engine, handle, B, R, P

    def read_kv(self, layer: int, which: str) -> np.ndarray:
        idx = {"k_ref": 0, "v_ref": 1, "k_text": 2, "v_text": 3}[which]
        n = self.R if idx < 2 else self.P
        out = np.empty((self.B, 8, n, 120), dtype=np.float32)
        _cabi.check(self._engine._lib.stts_cond_read_kv(self._engine._h, self._h, layer, idx, C.c_void_p(out.ctypes.data)),
                    self._engine._h)
        return out

    def free(self) -> None:
        if self._h is not None:
            self._engine._lib.stts_cond_free(self._engine._h, self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Engine:
    def __init__(self, device: int = 0, precision: str = "fast"):
        """precision: "fast" = bf16 GEMM / attention operands (the benchmarked build); "tight" = the parity build of the
        same kernels with fp16 operands (11-bit significand like TF32; csrc/op16.cuh).  fp32 accumulation, norms,
        softmax, RoPE and sampler in both."""
        self._lib = _cabi.lib(precision)
        self.precision = precision
        cfg = _cabi.Config(device=device)
        h = C.c_void_p()
        _cabi.check(self._lib.stts_create(C.byref(cfg), C.byref(h)), None)
        self._h = h
        self.device = device
        self._ready = False

    def clone(self) -> "Engine":
        """A second handle that shares this engine's weights (no copy) and has its own streams / plans: run it from
        another host thread to keep two batches in flight on the GPU (stts_engine_clone).  Close clones first."""
        h = C.c_void_p()
        _cabi.check(self._lib.stts_engine_clone(self._h, C.byref(h)), self._h)
        c = Engine.__new__(Engine)
        c._lib, c._h, c.device, c._ready, c._parent = self._lib, h, self.device, True, self  # parent outlives the clone
        c.precision = self.precision
        return c

    # ------------------------------------------------------------------ weights
    def load_state_dicts(self, dit_sd: Optional[Dict[str, "np.ndarray"]], vocoder_sd: Optional[Dict[str, "np.ndarray"]],
                         encoder_sd: Optional[Dict[str, "np.ndarray"]] = None) -> None:
        """fp32 tensors under the reference's own key names (DiTModel.state_dict(), HF decoder state_dict() and,
        optionally, HF encoder state_dict() for the clone path).  Any model may be None: an engine with only the
        codec models serves the standalone ``Decoder`` / ``Encoder`` (codec/onnx.py:34-75); operators of a model
        that was not loaded raise."""
        for model, sd in ((0, dit_sd or {}), (1, vocoder_sd or {}), (2, encoder_sd or {})):
            for name, t in sd.items():
                a = t.detach().cpu().numpy() if _is_torch(t) else np.asarray(t)
                a = np.require(a, dtype=np.float32, requirements=["C"])  # keeps 0-d tensors 0-d
                shape = (C.c_int64 * max(a.ndim, 1))(*a.shape)
                _cabi.check(self._lib.stts_load_weight(self._h, model, name.encode(), C.c_void_p(a.ctypes.data), a.ndim,
                                                       shape), self._h)
        _cabi.check(self._lib.stts_finalize_weights(self._h), self._h)
        self._ready = True

    # ------------------------------------------------------------------ operators
    def encode_conditions(self, ref, ref_len: Sequence[int], phonemes, ph_len: Sequence[int]) -> Conditions:
        B, R, _ = ref.shape
        P = phonemes.shape[1]
        rp, m1, k1 = _buf(ref, np.float32, "ref")
        pp, m2, k2 = _buf(phonemes, np.int64, "phonemes")
        if m1 != m2:
            raise ValueError("ref and phonemes must live in the same memory space")
        rl, pl = _i64(ref_len), _i64(ph_len)
        h = C.c_void_p()
        _cabi.check(self._lib.stts_encode_conditions(self._h, rp, C.c_void_p(rl.ctypes.data), pp,
                                                     C.c_void_p(pl.ctypes.data), B, R, P, m1, C.byref(h)), self._h)
        return Conditions(self, h, B, R, P)

    def denoise_step(self, cond: Conditions, x_t, frames: Sequence[int], t: Sequence[float]):
        B, T, _ = x_t.shape
        xp, mem, keep = _buf(x_t, np.float32, "x_t")
        fr = _i64(frames)
        tt = np.ascontiguousarray(np.asarray(list(t), dtype=np.float32))
        out, op = self._out_like(x_t, (B, T, LATENT_DIM), mem)
        _cabi.check(self._lib.stts_denoise_step(self._h, cond._h, xp, C.c_void_p(fr.ctypes.data),
                                                C.c_void_p(tt.ctypes.data), B, T, mem, op), self._h)
        return out

    def sample(self, cond: Conditions, frames: Sequence[int], T: int, noise=None, seed: int = 0, steps: int = 4,
               timesteps: Optional[Sequence[float]] = None, device_out: bool = False):
        B = cond.B
        fr = _i64(frames)
        ts = None if timesteps is None else np.ascontiguousarray(np.asarray(list(timesteps), dtype=np.float32))
        if noise is not None:
            np_, mem, keep = _buf(noise, np.float32, "noise")
            if tuple(noise.shape) != (steps, B, T, LATENT_DIM):
                raise ValueError(f"noise must be {(steps, B, T, LATENT_DIM)}, got {tuple(noise.shape)}")
        else:
            np_, mem = None, (_cabi.MEM_DEVICE if device_out else _cabi.MEM_HOST)
        out, op = self._out_like(noise, (B, T, LATENT_DIM), mem)
        _cabi.check(self._lib.stts_sample(self._h, cond._h, C.c_void_p(fr.ctypes.data), B, T, steps,
                                          None if ts is None else C.c_void_p(ts.ctypes.data), np_, seed, mem, op),
                    self._h)
        return out

    def encode_conditions_cfg(self, ref, ref_len: Sequence[int], phonemes, ph_len: Sequence[int]) -> Conditions:
        """Conditions of the 3-way classifier-free-guidance batch of the teacher (distill.py:74-96): rows
        [cond | text dropped | speaker dropped]; a dropped condition is a zero length (and zeroed ids)."""
        ref = np.asarray(ref.detach().cpu().numpy() if _is_torch(ref) else ref, dtype=np.float32)
        ids = np.asarray(phonemes.detach().cpu().numpy() if _is_torch(phonemes) else phonemes, dtype=np.int64)
        B = ref.shape[0]
        ref3 = np.concatenate([ref, ref, np.zeros_like(ref)], axis=0)
        ids3 = np.concatenate([ids, np.zeros_like(ids), ids], axis=0)
        rl, pl = list(map(int, ref_len)), list(map(int, ph_len))
        return self.encode_conditions(ref3, rl + rl + [0] * B, ids3, pl + [0] * B + pl)

    def sample_teacher(self, cond3: Conditions, frames: Sequence[int], T: int, steps: int = 128, cfg_text: float = 2.0,
                       cfg_speaker: float = 1.5, noise=None, seed: int = 0, device_out: bool = False):
        """DDIM + 3-way CFG teacher sampler (stts_sample_teacher).  noise: optional start x_1 of shape (B,T,64)."""
        if cond3.B % 3 != 0:
            raise ValueError("cond3 must come from encode_conditions_cfg (3*B rows)")
        B = cond3.B // 3
        fr = _i64(frames)
        if noise is not None:
            np_, mem, keep = _buf(noise, np.float32, "noise")
            if tuple(noise.shape) != (B, T, LATENT_DIM):
                raise ValueError(f"noise must be {(B, T, LATENT_DIM)}, got {tuple(noise.shape)}")
        else:
            np_, mem = None, (_cabi.MEM_DEVICE if device_out else _cabi.MEM_HOST)
        out, op = self._out_like(noise, (B, T, LATENT_DIM), mem)
        _cabi.check(self._lib.stts_sample_teacher(self._h, cond3._h, C.c_void_p(fr.ctypes.data), B, T, steps,
                                                  cfg_text, cfg_speaker, np_, seed, mem, op), self._h)
        return out

    def encode_audio(self, audio):
        """Codec encoder (codec/onnx.py:56-75): audio (B, N) or (B, 1, N) fp32 @ 24 kHz -> latents (B, N // 3200, 64).
        A tail shorter than one hop is dropped: the `transformers` port this engine is pinned on emits floor(N / 3200)
        latents and is causal, so no latent changes.  (The published encoder.onnx is an export of Microsoft's VibeVoice,
        whose non-streaming SConv1d may right-pad to ceil(N / 3200); unverified without the asset -- if it does, zero-pad
        the clip to a whole hop before calling this.)"""
        if audio.ndim == 3:
            audio = audio[:, 0]
        B, N = audio.shape
        n = (N // HOP_SIZE) * HOP_SIZE
        if n == 0:
            raise ValueError("audio shorter than one hop (3200 samples)")
        audio = audio[:, :n]
        ap, mem, keep = _buf(audio, np.float32, "audio")
        out, op = self._out_like(audio, (B, n // HOP_SIZE, LATENT_DIM), mem)
        _cabi.check(self._lib.stts_encode_audio(self._h, ap, B, n, mem, op), self._h)
        return out

    def resample(self, audio, sr_from: int, sr_to: int = 24_000):
        """``resample_hq`` (infer/utils.py:7-23) on the device: audio (B, N) fp32 at ``sr_from`` ->
        (B, ceil(N * sr_to / sr_from)).  numpy in -> numpy out; cuda tensor in -> cuda tensor out."""
        if audio.ndim != 2:
            raise ValueError(f"audio must be (B, N), got {tuple(audio.shape)}")
        B, N = audio.shape
        n_out = int(self._lib.stts_resample_length(N, int(sr_from), int(sr_to)))
        if n_out < 1:
            raise ValueError("bad resample arguments")
        ap, mem, keep = _buf(audio, np.float32, "audio")
        out, op = self._out_like(audio, (B, n_out), mem)
        _cabi.check(self._lib.stts_resample(self._h, ap, B, N, int(sr_from), int(sr_to), mem, op), self._h)
        return out

    def decode(self, latents):
        B, T, _ = latents.shape
        lp, mem, keep = _buf(latents, np.float32, "latents")
        out, op = self._out_like(latents, (B, T * HOP_SIZE), mem)
        _cabi.check(self._lib.stts_decode(self._h, lp, B, T, mem, op), self._h)
        return out

    def synthesize(self, ref, ref_len, phonemes, ph_len, frames, T: int, noise=None, seed: int = 0, steps: int = 4,
                   timesteps: Optional[Sequence[float]] = None, out=None):
        """Padded-batch SmallTTS.synthesize: returns audio [B, T*3200] (numpy if inputs are numpy, torch cuda if
        inputs are cuda tensors).  ``out`` may be a preallocated (pinned) numpy array / cuda tensor."""
        B, R, _ = ref.shape
        P = phonemes.shape[1]
        rp, mem, k1 = _buf(ref, np.float32, "ref")
        pp, m2, k2 = _buf(phonemes, np.int64, "phonemes")
        np_ = None
        if noise is not None:
            np_, m3, k3 = _buf(noise, np.float32, "noise")
            if tuple(noise.shape) != (steps, B, T, LATENT_DIM):
                raise ValueError(f"noise must be {(steps, B, T, LATENT_DIM)}, got {tuple(noise.shape)}")
            if m3 != mem:
                raise ValueError("all tensors must live in the same memory space")
        if m2 != mem:
            raise ValueError("all tensors must live in the same memory space")
        rl, pl, fr = _i64(ref_len), _i64(ph_len), _i64(frames)
        ts = None if timesteps is None else np.ascontiguousarray(np.asarray(list(timesteps), dtype=np.float32))
        if out is None:
            out, op = self._out_like(ref, (B, T * HOP_SIZE), mem)
        else:
            # the engine writes B*T*3200 floats straight through this pointer: never accept a buffer it could overrun
            # and never let a silent copy (non-contiguous / wrong dtype) swallow the result
            ok_dtype = str(out.dtype).endswith("float32")
            contiguous = out.is_contiguous() if _is_torch(out) else bool(out.flags["C_CONTIGUOUS"])
            if tuple(out.shape) != (B, T * HOP_SIZE) or not ok_dtype or not contiguous:
                raise ValueError(f"out must be a C-contiguous float32 buffer of shape {(B, T * HOP_SIZE)}, got "
                                 f"{tuple(out.shape)} {out.dtype}")
            op, mo, _ = _buf(out, np.float32, "out")
            if mo != mem:
                raise ValueError("out must live in the same memory space as the inputs")
        _cabi.check(self._lib.stts_synthesize(self._h, rp, C.c_void_p(rl.ctypes.data), pp, C.c_void_p(pl.ctypes.data),
                                              C.c_void_p(fr.ctypes.data), B, R, P, T, steps,
                                              None if ts is None else C.c_void_p(ts.ctypes.data), np_, seed, mem, op),
                    self._h)
        return out

    # ------------------------------------------------------------------ misc
    def timings(self) -> Dict[str, float]:
        t = _cabi.Timing()
        _cabi.check(self._lib.stts_get_timings(self._h, C.byref(t)), self._h)
        return {n: getattr(t, n) for n, _ in _cabi.Timing._fields_}

    def vocoder_ms(self) -> Dict[str, float]:
        return {"tail_hbm": self._lib.stts_last_vocoder_ms(self._h, 0),
                "front_tensor": self._lib.stts_last_vocoder_ms(self._h, 1)}

    def timer_start(self) -> None:
        _cabi.check(self._lib.stts_timer_start(self._h), self._h)

    def timer_stop(self) -> float:
        ms = C.c_float()
        _cabi.check(self._lib.stts_timer_stop(self._h, C.byref(ms)), self._h)
        return float(ms.value)

    def pinned(self, shape, dtype=np.float32) -> np.ndarray:
        """numpy array backed by page-locked host memory (cudaHostAlloc) for truly asynchronous copies."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = self._lib.stts_host_alloc(max(n, 1))
        if not p:
            raise MemoryError("cudaHostAlloc failed")
        buf = (C.c_byte * max(n, 1)).from_address(p)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    @staticmethod
    def launch_count() -> int:
        """Kernels launched so far by every loaded build of the library in this process."""
        return sum(int(l.stts_launch_count()) for l in _cabi._libs.values()) if _cabi._libs else int(_cabi.lib().stts_launch_count())

    def _out_like(self, like, shape, mem):
        if mem == _cabi.MEM_DEVICE:
            import torch

            out = torch.empty(shape, dtype=torch.float32, device=f"cuda:{self.device}")
            # the caching allocator may hand out memory whose previous user is still running on torch's stream
            torch.cuda.current_stream(out.device).synchronize()
            return out, C.c_void_p(out.data_ptr())
        out = np.empty(shape, dtype=np.float32)
        return out, C.c_void_p(out.ctypes.data)

    def close(self) -> None:
        if self._h is not None:
            self._lib.stts_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pad_batch(ref_list, ids_list: Sequence[Sequence[int]], frames: Sequence[int], r_mult: int = 1, p_mult: int = 1):
    """Ragged python inputs -> padded numpy batch (ref [B,R,64], ref_len, ids [B,P], ph_len).  ``r_mult`` / ``p_mult``
    round the padded widths up to a multiple (the padding is masked by the lengths): fewer distinct shapes, so the
    engine re-uses its per-shape plans / CUDA graphs instead of building new ones."""
    B = len(ref_list)
    refs = [np.asarray(r.detach().cpu().numpy() if _is_torch(r) else r, dtype=np.float32) for r in ref_list]
    R = max(r.shape[0] for r in refs)
    P = max(1, max(len(p) for p in ids_list))
    R, P = -(-R // r_mult) * r_mult, -(-P // p_mult) * p_mult
    ref = np.zeros((B, R, LATENT_DIM), dtype=np.float32)
    ids = np.zeros((B, P), dtype=np.int64)
    for i in range(B):
        if refs[i].ndim != 2 or refs[i].shape[1] != LATENT_DIM:
            raise ValueError(f"ref_latents[{i}] must be (R, 64), got {refs[i].shape}")
        ref[i, : refs[i].shape[0]] = refs[i]
        ids[i, : len(ids_list[i])] = np.asarray(list(ids_list[i]), dtype=np.int64)
    return ref, [r.shape[0] for r in refs], ids, [len(p) for p in ids_list]

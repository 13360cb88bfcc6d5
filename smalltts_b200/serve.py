"""Serving shim on the B200 engine: the reference server's ``Pipeline`` (src/server/src/pipeline.rs:40-112) and a
request micro-batcher in place of its ``Arc<Mutex<Pipeline>>`` (src/server/src/main.rs:25,138-147).

The reference serialises requests behind one mutex and runs each as batch 1 (codec encode -> condition encode ->
4 denoiser steps -> codec decode, ``synthesize_timed``).  On a B200 a single 10 s request leaves the GPU mostly idle
(the DiT is latency-bound below ~8 utterances), so the shim gathers the requests that arrive within a
short window into ONE ragged engine call: reference clips are right-padded and encoded together (the codec encoder is
causal, so a clip's latents do not depend on the padding), prompts are padded to the longest and masked by length.

Out of scope here, as in DESIGN.md: x402 payments, CORS, multipart parsing of arbitrary audio containers (the HTTP
front below takes 16-bit PCM / float32 WAV only) and the espeak phonemizer (token ids can be posted directly; text
needs the reference's ``smalltts.data.phonemization`` package to be importable).
"""
from __future__ import annotations

import io
import json
import math
import queue
import struct
import threading
import time
import wave
from concurrent.futures import Future
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

SAMPLE_RATE = 24_000  # pipeline.rs:11
HOP_SIZE = 3_200  # pipeline.rs:12
ENGINE_MAX_LEN = 4096  # frames / reference hops / tokens one engine call accepts (RoPE table size, dit.py:139)


@dataclass
class Timing:
    """pipeline.rs:29-37 (milliseconds).  Device-timed by the engine (CUDA events), total = wall clock of the call."""

    codec_enc_ms: float = 0.0
    cond_enc_ms: float = 0.0
    denoise_ms: float = 0.0
    codec_dec_ms: float = 0.0
    total_ms: float = 0.0
    batch: int = 1  # how many requests shared this engine call (always 1 in the reference)


def seq_len_for(duration_sec: float) -> int:
    """pipeline.rs:71: ``ceil(duration * 24000 / 3200).max(1)`` -- the server rounds UP where infer/onnx.py:84 floors."""
    return max(1, int(math.ceil(duration_sec * SAMPLE_RATE / HOP_SIZE)))


@dataclass
class Request:
    ref_audio: np.ndarray  # mono fp32 @ 24 kHz
    token_ids: List[int]
    duration_sec: float
    future: Future = field(default_factory=Future)
    t_submit: float = field(default_factory=time.perf_counter)


class Pipeline:
    """``Pipeline::load`` / ``synthesize`` / ``synthesize_timed`` (pipeline.rs:40-112) for one request or a batch.

    ``tts`` is a :class:`smalltts_b200.infer.SmallTTS` whose engine carries the codec encoder (the server always
    encodes the posted reference audio, pipeline.rs:74-76)."""

    def __init__(self, tts, shape_buckets: Optional[Sequence[int]] = None) -> None:
        self.tts = tts
        # Requests come in every shape.  Rounding the padded (R, P, T) up lets the engine's per-shape plans (buffers +
        # CUDA graphs, an LRU of 6) be re-used instead of rebuilt; T costs vocoder work, hence the small multiple in
        # e.g. (8, 16, 5).  Off by default; the buckets live on the Pipeline (the caller's SmallTTS is not modified)
        # and are applied per call.
        self.shape_buckets = None if shape_buckets is None else tuple(int(x) for x in shape_buckets)

    @classmethod
    def load(cls, cond_encoder_path: str = "assets/dmd/condition_encoder.onnx",
             denoiser_path: str = "assets/dmd/denoiser.onnx", codec_decoder_path: str = "assets/codec/decoder.onnx",
             codec_encoder_path: str = "assets/codec/encoder.onnx", device: int = 0) -> "Pipeline":
        from .infer import SmallTTS

        return cls(SmallTTS(cond_encoder_path, denoiser_path, codec_decoder_path,
                            codec_encoder_path=codec_encoder_path, device=device))

    def synthesize(self, ref_audio: np.ndarray, token_ids: Sequence[int], duration_sec: float) -> np.ndarray:
        return self.synthesize_timed(ref_audio, token_ids, duration_sec)[0]

    def synthesize_timed(self, ref_audio: np.ndarray, token_ids: Sequence[int], duration_sec: float
                         ) -> Tuple[np.ndarray, Timing]:
        audio, timing = self.synthesize_many([ref_audio], [token_ids], [duration_sec])
        return audio[0], timing

    def synthesize_many(self, ref_audios: Sequence[np.ndarray], token_ids: Sequence[Sequence[int]],
                        durations: Sequence[float]) -> Tuple[List[np.ndarray], Timing]:
        """One engine pass for several requests: returns [(samples_i,)] fp32 and the stage timing of the pass."""
        t0 = time.perf_counter()
        eng = self.tts.engine
        n_hops = [len(a) // HOP_SIZE for a in ref_audios]
        if min(n_hops) < 1:
            raise ValueError("reference audio shorter than one codec hop (3200 samples at 24 kHz)")
        padded = np.zeros((len(ref_audios), max(n_hops) * HOP_SIZE), dtype=np.float32)
        for i, a in enumerate(ref_audios):
            padded[i, : n_hops[i] * HOP_SIZE] = np.asarray(a, dtype=np.float32).reshape(-1)[: n_hops[i] * HOP_SIZE]
        lat = eng.encode_audio(padded)  # causal encoder: row i's first n_hops[i] latents ignore the padding
        codec_enc_ms = eng.timings()["codec_enc_ms"]
        refs = [lat[i, : n_hops[i]] for i in range(len(ref_audios))]
        # the server's frame count rounds up (pipeline.rs:71); SmallTTS.synthesize_batch floors durations
        frames = [seq_len_for(d) for d in durations]
        durs = [(f + 0.5) * HOP_SIZE / SAMPLE_RATE for f in frames]
        audio = self.tts.synthesize_batch(refs, [list(map(int, t)) for t in token_ids], durs,
                                          shape_buckets=self.shape_buckets)
        tm = eng.timings()
        timing = Timing(codec_enc_ms, tm["cond_enc_ms"], tm["denoise_ms"], tm["codec_dec_ms"],
                        (time.perf_counter() - t0) * 1e3, len(ref_audios))
        return [a.reshape(-1) for a in audio], timing


class MicroBatcher:
    """Replaces ``Arc<Mutex<Pipeline>>``: any number of request threads ``submit``; one worker thread owns the engine
    (the C-ABI handle is not thread-safe) and runs up to ``max_batch`` queued requests per engine pass, waiting at most
    ``max_wait_ms`` after the first request of a batch for more to arrive.  ``max_frames`` bounds the padded work of a
    batch (sum over requests of the longest duration in frames) so one long prompt does not inflate many short ones.

    ``run_batch`` may be a list of callables, one per engine replica (each replica = its own handle, stream and worker
    thread; replicas may share a GPU): batches are formed one at a time but run concurrently.  Two batches of 8 x 10 s
    in flight on one B200 give 1.17x the throughput of one (profiles/r01_concurrent_replicas.txt): the second batch
    fills the SMs that the latency-bound DiT GEMMs of the first leave idle."""

    def __init__(self, run_batch, max_batch: int = 8, max_wait_ms: float = 2.0, max_frames: int = 8 * 75) -> None:
        if max_batch < 1:
            raise ValueError("max_batch must be >= 1")
        self._runs = list(run_batch) if isinstance(run_batch, (list, tuple)) else [run_batch]
        if not self._runs:
            raise ValueError("run_batch must not be empty")
        self.max_batch, self.max_wait_s, self.max_frames = max_batch, max_wait_ms / 1e3, max_frames
        self._q: "queue.Queue[Optional[Request]]" = queue.Queue()
        self._carry: Optional[Request] = None
        self.batches_run = 0
        self.requests_run = 0
        self._closed = False
        self._form = threading.Lock()  # one worker at a time forms a batch (and touches _carry)
        self._stats = threading.Lock()
        self._workers = [threading.Thread(target=self._loop, args=(fn,), name=f"stts-microbatcher-{k}", daemon=True)
                         for k, fn in enumerate(self._runs)]
        for w in self._workers:
            w.start()

    # ------------------------------------------------------------------ client side
    def submit(self, ref_audio: np.ndarray, token_ids: Sequence[int], duration_sec: float) -> Future:
        """-> Future of (audio fp32 (samples,), Timing).  Validation errors surface on the future, like the HTTP 400s
        of main.rs:124-131."""
        req = Request(np.asarray(ref_audio, dtype=np.float32).reshape(-1), list(map(int, token_ids)), float(duration_sec))
        if self._closed:
            req.future.set_exception(RuntimeError("batcher is closed"))
        elif not (req.duration_sec > 0) or not math.isfinite(req.duration_sec):
            req.future.set_exception(ValueError("duration must be a positive number of seconds"))
        elif len(req.ref_audio) < HOP_SIZE:
            req.future.set_exception(ValueError("reference audio shorter than one codec hop (3200 samples at 24 kHz)"))
        elif not (1 <= len(req.token_ids) <= ENGINE_MAX_LEN):
            req.future.set_exception(ValueError(f"between 1 and {ENGINE_MAX_LEN} phoneme tokens are required"))
        elif len(req.ref_audio) // HOP_SIZE > ENGINE_MAX_LEN or seq_len_for(req.duration_sec) > ENGINE_MAX_LEN:
            req.future.set_exception(ValueError(f"reference audio and duration are limited to {ENGINE_MAX_LEN} codec frames"))
        else:
            self._q.put(req)
        return req.future

    def synthesize(self, ref_audio: np.ndarray, token_ids: Sequence[int], duration_sec: float,
                   timeout: Optional[float] = None) -> np.ndarray:
        return self.submit(ref_audio, token_ids, duration_sec).result(timeout)[0]

    def close(self) -> None:
        if not self._closed:
            self._closed = True
            self._q.put(None)
            for w in self._workers:
                w.join()
            while True:  # whatever was queued behind the shutdown marker
                try:
                    r = self._q.get_nowait()
                except queue.Empty:
                    break
                if r is not None:
                    r.future.set_exception(RuntimeError("batcher is closed"))

    # ------------------------------------------------------------------ worker
    def _take_batch(self) -> Optional[List[Request]]:
        first = self._carry if self._carry is not None else self._q.get()
        self._carry = None
        if first is None:
            self._q.put(None)  # pass the shutdown marker on to the other workers
            return None
        batch = [first]
        longest = seq_len_for(first.duration_sec)
        deadline = time.perf_counter() + self.max_wait_s
        while len(batch) < self.max_batch:
            remaining = deadline - time.perf_counter()
            try:
                nxt = self._q.get(timeout=remaining) if remaining > 0 else self._q.get_nowait()
            except queue.Empty:
                break
            if nxt is None:
                self._q.put(None)  # leave the shutdown marker for the next round
                break
            cand = max(longest, seq_len_for(nxt.duration_sec))
            if cand * (len(batch) + 1) > max(self.max_frames, cand):
                self._carry = nxt  # would pad too much: it opens the next batch
                break
            batch.append(nxt)
            longest = cand
        return batch

    def _loop(self, run) -> None:
        while True:
            with self._form:
                batch = self._take_batch()
            if batch is None:
                break
            self._run_batch(run, batch)

    @staticmethod
    def _resolve(fut: Future, result=None, exc: Optional[BaseException] = None) -> None:
        """A client may have cancelled its future; never let that take the worker thread down."""
        try:
            if fut.done():
                return
            if exc is not None:
                fut.set_exception(exc)
            else:
                fut.set_result(result)
        except Exception:  # noqa: BLE001 - InvalidStateError from a race with cancel()
            pass

    def _run_batch(self, run, batch: List[Request]) -> None:
        try:
            audio, timing = run([r.ref_audio for r in batch], [r.token_ids for r in batch],
                                [r.duration_sec for r in batch])
        except Exception as exc:  # noqa: BLE001
            if len(batch) == 1:  # this request is the offender (HTTP 400/500 for it alone)
                self._resolve(batch[0].future, exc=exc)
            else:  # the reference serves requests independently: retry one at a time so only the offender fails
                for r in batch:
                    self._run_batch(run, [r])
            return
        for r, a in zip(batch, audio):
            self._resolve(r.future, result=(a, timing))
        with self._stats:  # not _form: an idle worker holds that one while it waits for the next request
            self.batches_run += 1
            self.requests_run += len(batch)


# ---------------------------------------------------------------------------------------------- WAV + HTTP front
def decode_wav(data: bytes) -> Tuple[np.ndarray, int]:
    """16-bit PCM or 32-bit float WAV bytes -> (mono fp32, sample_rate).  (audio.rs decodes any container with
    symphonia; this front keeps to what the stdlib can parse.)"""
    if len(data) >= 22 and data[:4] == b"RIFF" and struct.unpack_from("<H", data, 20)[0] == 3:  # IEEE float
        pos = 12
        fmt, pcm = None, None
        while pos + 8 <= len(data):
            cid, size = data[pos : pos + 4], struct.unpack_from("<I", data, pos + 4)[0]
            if cid == b"fmt ":
                fmt = struct.unpack_from("<HHIIHH", data, pos + 8)
            elif cid == b"data":
                pcm = data[pos + 8 : pos + 8 + size]
            pos += 8 + size + (size & 1)
        if fmt is None or pcm is None or fmt[5] != 32:
            raise ValueError("unsupported float WAV")
        x = np.frombuffer(pcm, dtype="<f4").reshape(-1, fmt[1])
        return x.mean(axis=1).astype(np.float32), int(fmt[2])
    with wave.open(io.BytesIO(data), "rb") as w:
        if w.getsampwidth() != 2:
            raise ValueError("only 16-bit PCM and 32-bit float WAV are supported")
        x = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").reshape(-1, w.getnchannels())
        return (x.astype(np.float32) / 32768.0).mean(axis=1), w.getframerate()


def encode_wav(audio: np.ndarray, sample_rate: int = SAMPLE_RATE) -> bytes:
    """fp32 [-1, 1] -> 16-bit PCM WAV bytes (audio.rs encode_wav / the scripts' ``subtype="PCM_16"``)."""
    pcm = (np.clip(np.asarray(audio, dtype=np.float32).reshape(-1), -1.0, 1.0) * 32767.0).round().astype("<i2")
    buf = io.BytesIO()
    with wave.open(buf, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(sample_rate)
        w.writeframes(pcm.tobytes())
    return buf.getvalue()


def _multipart(body: bytes, content_type: str) -> dict:
    """Minimal multipart/form-data reader: {field name: bytes}."""
    marker = "boundary="
    if marker not in content_type:
        raise ValueError("multipart boundary missing")
    boundary = ("--" + content_type.split(marker, 1)[1].strip().strip('"')).encode()
    out = {}
    for part in body.split(boundary)[1:]:
        if part.startswith(b"--"):
            break
        head, _, payload = part.lstrip(b"\r\n").partition(b"\r\n\r\n")
        name = None
        for line in head.decode(errors="replace").split("\r\n"):
            if line.lower().startswith("content-disposition") and 'name="' in line:
                name = line.split('name="', 1)[1].split('"', 1)[0]
        if name is not None:
            out[name] = payload[:-2] if payload.endswith(b"\r\n") else payload
    return out


def make_handler(batcher: MicroBatcher, resample: Callable[[np.ndarray, int], np.ndarray],
                 phonemize: Optional[Callable[[str], List[int]]] = None):
    """http.server handler with the routes of main.rs:57-60: GET /health, POST /synthesize?duration=<s> (multipart
    fields ``audio`` = WAV and ``text``, or ``tokens`` = JSON list of ids)."""
    from http.server import BaseHTTPRequestHandler
    from urllib.parse import parse_qs, urlparse

    class Handler(BaseHTTPRequestHandler):
        def _send(self, code: int, body: bytes, ctype: str = "text/plain") -> None:
            self.send_response(code)
            self.send_header("content-type", ctype)
            self.send_header("content-length", str(len(body)))
            self.end_headers()
            self.wfile.write(body)

        def log_message(self, fmt, *args):  # quiet
            pass

        def do_GET(self):  # noqa: N802
            if urlparse(self.path).path == "/health":
                self._send(200, b"ok")
            else:
                self._send(404, b"not found")

        def do_POST(self):  # noqa: N802
            url = urlparse(self.path)
            if url.path != "/synthesize":
                return self._send(404, b"not found")
            try:
                duration = float(parse_qs(url.query)["duration"][0])
                n = int(self.headers.get("content-length", "0"))
                if n > 2 * 1024 * 1024:  # main.rs:82 RequestBodyLimitLayer
                    return self._send(413, b"body too large")
                fields = _multipart(self.rfile.read(n), self.headers.get("content-type", ""))
                if "audio" not in fields:
                    return self._send(400, b"missing 'audio'")
                if "tokens" in fields:
                    tokens = [int(t) for t in json.loads(fields["tokens"].decode())]
                elif "text" in fields:
                    if phonemize is None:
                        return self._send(500, b"phonemize failed: no phonemizer installed; post 'tokens'")
                    tokens = phonemize(fields["text"].decode())
                else:
                    return self._send(400, b"missing 'text'")
                wav, sr = decode_wav(fields["audio"])
                ref = resample(wav, sr)
            except (KeyError, ValueError, wave.Error, EOFError, struct.error) as exc:
                return self._send(400, f"bad request: {exc}".encode())
            try:
                audio, _ = batcher.submit(ref, tokens, duration).result()
            except ValueError as exc:
                return self._send(400, f"bad request: {exc}".encode())
            except Exception as exc:
                return self._send(500, f"inference failed: {exc}".encode())
            self._send(200, encode_wav(audio), "audio/wav")

    return Handler


def serve(pipeline, host: str = "0.0.0.0", port: int = 3000, max_batch: int = 8, max_wait_ms: float = 2.0):
    """Blocking HTTP server (main.rs:84-88).  ``pipeline``: a :class:`Pipeline` or a list of replicas (same or different
    GPUs) that run batches concurrently.  ``PORT`` handling and the payment layer stay with the deployment."""
    from http.server import ThreadingHTTPServer

    pipelines = list(pipeline) if isinstance(pipeline, (list, tuple)) else [pipeline]
    pipeline = pipelines[0]
    batcher = MicroBatcher([p.synthesize_many for p in pipelines], max_batch=max_batch, max_wait_ms=max_wait_ms)
    # Request threads resample before they queue; they use a second engine handle on the same device (resampling needs
    # no weights) so that the batcher's worker stays the only user of the pipeline's handle.
    from .engine import Engine

    rs_engine, rs_lock = Engine(pipeline.tts.engine.device), threading.Lock()

    def resample(wav: np.ndarray, sr: int) -> np.ndarray:
        if sr == SAMPLE_RATE:
            return wav
        with rs_lock:
            return rs_engine.resample(wav[None], sr, SAMPLE_RATE)[0]

    try:
        from smalltts.data.phonemization.phonemes import get_token_ids as phonemize  # the reference's front-end
    except Exception:
        phonemize = None
    httpd = ThreadingHTTPServer((host, port), make_handler(batcher, resample, phonemize))
    try:
        httpd.serve_forever()
    finally:
        batcher.close()
        httpd.server_close()

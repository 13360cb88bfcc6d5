"""Serving shim (smalltts_b200/serve.py): micro-batcher semantics, WAV helpers and the HTTP front, on a fake
pipeline -- the engine itself is covered by the GPU tests."""
import json
import threading
import time
import urllib.error
import urllib.request

import numpy as np
import pytest

from smalltts_b200 import serve


class FakePipeline:
    """Stands in for Pipeline.synthesize_many: request i gets seq_len*3200 samples filled with its first token id."""

    def __init__(self, delay=0.0, fail_on=None):
        self.calls = []
        self.delay, self.fail_on = delay, fail_on
        self.threads = set()

    def __call__(self, refs, tokens, durations):
        self.threads.add(threading.get_ident())
        self.calls.append((len(refs), list(durations)))
        time.sleep(self.delay)
        if self.fail_on is not None and any(t and t[0] == self.fail_on for t in tokens):
            raise RuntimeError("inference failed: boom")
        out = [np.full(serve.seq_len_for(d) * 3200, float(t[0] if t else 0), np.float32) for t, d in zip(tokens, durations)]
        return out, serve.Timing(1, 2, 3, 4, 10, len(refs))


REF = np.zeros(3200, np.float32)


def test_seq_len_rounds_up_like_the_server():
    assert serve.seq_len_for(2.0) == 15 and serve.seq_len_for(2.01) == 16 and serve.seq_len_for(1e-6) == 1  # pipeline.rs:71


def test_requests_arriving_together_share_one_pass():
    fake = FakePipeline(delay=0.05)
    b = serve.MicroBatcher(fake, max_batch=4, max_wait_ms=200)
    try:
        futs = [b.submit(REF, [i + 1], 1.0) for i in range(6)]
        res = [f.result(10) for f in futs]
    finally:
        b.close()
    for i, (audio, timing) in enumerate(res):  # every request gets its own row back
        assert audio.shape == (8 * 3200,) and audio[0] == i + 1
    assert [c[0] for c in fake.calls] == [4, 2]
    assert res[0][1].batch == 4 and res[5][1].batch == 2
    assert len(fake.threads) == 1  # one worker thread owns the engine
    assert b.batches_run == 2 and b.requests_run == 6


def test_single_request_does_not_wait_for_a_full_batch():
    fake = FakePipeline()
    b = serve.MicroBatcher(fake, max_batch=8, max_wait_ms=30)
    try:
        t0 = time.perf_counter()
        audio = b.synthesize(REF, [7], 0.5, timeout=10)
        dt = time.perf_counter() - t0
    finally:
        b.close()
    assert audio[0] == 7 and fake.calls == [(1, [0.5])] and dt < 2.0


def test_padding_budget_splits_long_and_short_prompts():
    fake = FakePipeline(delay=0.02)
    b = serve.MicroBatcher(fake, max_batch=8, max_wait_ms=200, max_frames=100)
    try:
        futs = [b.submit(REF, [1], 1.0), b.submit(REF, [2], 1.0), b.submit(REF, [3], 10.0), b.submit(REF, [4], 1.0)]
        [f.result(10) for f in futs]
    finally:
        b.close()
    # 8+8 frames fit; adding the 75-frame prompt would pad 3 rows to 75 (225 > 100): it opens the next pass, where
    # the following 1 s prompt would again be padded to 75 frames (150 > 100)
    assert [c[0] for c in fake.calls] == [2, 1, 1]


def test_errors_reach_the_right_futures_and_the_worker_survives():
    fake = FakePipeline(fail_on=13)
    b = serve.MicroBatcher(fake, max_batch=1, max_wait_ms=1)
    try:
        bad, good = b.submit(REF, [13], 1.0), b.submit(REF, [5], 1.0)
        with pytest.raises(RuntimeError, match="boom"):
            bad.result(10)
        assert good.result(10)[0][0] == 5
        with pytest.raises(ValueError):
            b.submit(REF, [1], 0.0).result(1)
        with pytest.raises(ValueError):
            b.submit(REF[:100], [1], 1.0).result(1)
    finally:
        b.close()
    with pytest.raises(RuntimeError, match="closed"):
        b.submit(REF, [1], 1.0).result(1)


def test_one_bad_request_does_not_fail_its_batch_mates():
    """The reference serves requests independently (main.rs:138-147): when a batched pass fails, its requests are
    retried one at a time so only the offender gets the error; limits of the engine are checked at submit()."""
    fake = FakePipeline(delay=0.02, fail_on=13)
    b = serve.MicroBatcher(fake, max_batch=4, max_wait_ms=200)
    try:
        futs = [b.submit(REF, [t], 1.0) for t in (5, 13, 7)]
        assert futs[0].result(10)[0][0] == 5 and futs[2].result(10)[0][0] == 7
        with pytest.raises(RuntimeError, match="boom"):
            futs[1].result(10)
        assert [c[0] for c in fake.calls] == [3, 1, 1, 1]
        with pytest.raises(ValueError, match="tokens"):
            b.submit(REF, list(range(serve.ENGINE_MAX_LEN + 1)), 1.0).result(1)
        with pytest.raises(ValueError, match="tokens"):
            b.submit(REF, [], 1.0).result(1)
        with pytest.raises(ValueError, match="codec frames"):
            b.submit(REF, [1], 4096 * 3200 / 24000 + 1).result(1)
    finally:
        b.close()


def test_cancelled_future_does_not_kill_the_worker():
    fake = FakePipeline(delay=0.1)
    b = serve.MicroBatcher(fake, max_batch=1, max_wait_ms=1)
    try:
        first = b.submit(REF, [1], 1.0)
        second = b.submit(REF, [2], 1.0)
        assert second.cancel()  # still queued behind the first pass
        assert first.result(10)[0][0] == 1
        assert b.submit(REF, [3], 1.0).result(10)[0][0] == 3  # the worker thread is still alive
    finally:
        b.close()


def test_pipeline_keeps_buckets_to_itself():
    class T:  # SmallTTS stand-in
        shape_buckets = None

    t = T()
    p = serve.Pipeline(t)
    assert p.shape_buckets is None and t.shape_buckets is None  # bucketing is off by default
    p = serve.Pipeline(t, shape_buckets=(8, 16, 5))
    assert p.shape_buckets == (8, 16, 5) and t.shape_buckets is None  # the caller's object is not modified


def test_wav_roundtrip_pcm16_and_float():
    x = (0.5 * np.sin(np.arange(2400) * 0.05)).astype(np.float32)
    data = serve.encode_wav(x)
    y, sr = serve.decode_wav(data)
    assert sr == 24000 and y.shape == x.shape and np.abs(y - x).max() < 1e-4
    import struct

    stereo = np.stack([x, -x * 0.5], axis=1).astype("<f4")
    body = stereo.tobytes()
    hdr = b"RIFF" + struct.pack("<I", 36 + len(body)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 3, 2, 44100, 44100 * 8, 8, 32)
    y, sr = serve.decode_wav(hdr + b"data" + struct.pack("<I", len(body)) + body)
    assert sr == 44100 and np.allclose(y, 0.25 * x, atol=1e-7)
    with pytest.raises(Exception):
        serve.decode_wav(b"not a wav file at all")


def _post(url, fields):
    boundary = "XbOuNdArYx"
    body = b""
    for name, val in fields.items():
        body += f'--{boundary}\r\nContent-Disposition: form-data; name="{name}"\r\n\r\n'.encode() + val + b"\r\n"
    body += f"--{boundary}--\r\n".encode()
    req = urllib.request.Request(url, data=body, headers={"content-type": f"multipart/form-data; boundary={boundary}"})
    return urllib.request.urlopen(req, timeout=10)


def test_http_front_routes_and_status_codes():
    from http.server import ThreadingHTTPServer

    fake = FakePipeline()
    b = serve.MicroBatcher(fake, max_batch=2, max_wait_ms=1)
    seen = {}

    def resample(wav, sr):
        seen["sr"] = sr
        return wav

    httpd = ThreadingHTTPServer(("127.0.0.1", 0), serve.make_handler(b, resample, phonemize=lambda text: [len(text)]))
    port = httpd.server_address[1]
    th = threading.Thread(target=httpd.serve_forever, daemon=True)
    th.start()
    base = f"http://127.0.0.1:{port}"
    try:
        assert urllib.request.urlopen(base + "/health", timeout=10).read() == b"ok"
        wav = serve.encode_wav(np.zeros(4800, np.float32), 16000)
        r = _post(base + "/synthesize?duration=1.0", {"audio": wav, "tokens": json.dumps([9, 8]).encode()})
        assert r.status == 200 and r.headers["content-type"] == "audio/wav"
        y, sr = serve.decode_wav(r.read())
        assert sr == 24000 and y.shape == (8 * 3200,) and seen["sr"] == 16000
        r = _post(base + "/synthesize?duration=0.4", {"audio": wav, "text": b"hello"})  # text -> phonemizer
        assert serve.decode_wav(r.read())[0].shape == (3 * 3200,)
        for fields, query, code in (({"text": b"x"}, "duration=1", 400), ({"audio": wav}, "duration=1", 400),
                                    ({"audio": wav, "text": b"x"}, "", 400), ({"audio": b"junk", "text": b"x"}, "duration=1", 400)):
            with pytest.raises(urllib.error.HTTPError) as ei:
                _post(base + "/synthesize?" + query, fields)
            assert ei.value.code == code
    finally:
        httpd.shutdown()
        httpd.server_close()
        b.close()


def test_replicas_run_batches_concurrently_and_shut_down_cleanly():
    """One worker thread per engine replica: two batches are in flight at the same time, every request is served
    exactly once, close() stops every worker."""
    gate = threading.Barrier(2, timeout=5)
    fakes = [FakePipeline(), FakePipeline()]

    def make(f):
        def run(refs, tokens, durations):
            gate.wait()  # only passes if BOTH replicas are inside a pass at the same time
            return f(refs, tokens, durations)

        return run

    b = serve.MicroBatcher([make(f) for f in fakes], max_batch=2, max_wait_ms=20)
    try:
        futs = [b.submit(REF, [i + 1], 1.0) for i in range(4)]
        res = [f.result(10) for f in futs]
    finally:
        b.close()
    assert [int(a[0]) for a, _ in res] == [1, 2, 3, 4]
    assert sorted(c[0] for f in fakes for c in f.calls) == [2, 2] and all(len(f.calls) == 1 for f in fakes)
    assert b.batches_run == 2 and b.requests_run == 4
    assert not any(w.is_alive() for w in b._workers)
    with pytest.raises(ValueError):
        serve.MicroBatcher([])

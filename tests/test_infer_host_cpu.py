"""Host logic of the drop-in ``SmallTTS`` (smalltts_b200/infer.py) on a fake engine: padding, frame counts, seeds, noise
shapes, the list API and the ``devices=[...]`` split.  The engine itself is covered by the GPU tests."""
import numpy as np
import pytest

from smalltts_b200 import infer


class FakeEngine:
    instances = []

    def __init__(self, device=0, precision="fast"):
        self.device, self.precision = device, precision
        self.calls = []
        FakeEngine.instances.append(self)

    def load_state_dicts(self, *sds):
        self.loaded = [None if sd is None else len(sd) for sd in sds]

    def synthesize(self, ref, ref_len, ids, ph_len, frames, T, noise=None, seed=0, steps=4, timesteps=None, out=None):
        B = ref.shape[0]
        assert ref.shape[0] == B and ref.shape[1] >= max(ref_len) and ref.shape[2] == 64
        assert ids.shape[0] == B and ids.shape[1] >= max(1, max(ph_len))
        self.last_shape = (ref.shape[1], ids.shape[1])
        assert T >= max(frames) and all(1 <= f <= T for f in frames)
        if noise is not None:
            assert noise.shape == (steps, B, T, 64)
        self.calls.append(dict(B=B, T=T, frames=list(frames), seed=seed, steps=steps, noise=noise is not None,
                               ref_len=list(ref_len), ph_len=list(ph_len)))
        audio = np.zeros((B, T * 3200), np.float32)
        for b in range(B):  # encode (first token id, device) so that routing mistakes are visible
            audio[b, : frames[b] * 3200] = ids[b, 0] + 1000 * self.device
        return audio

    def close(self):
        pass


@pytest.fixture()
def fake(monkeypatch):
    FakeEngine.instances = []
    monkeypatch.setattr(infer, "Engine", FakeEngine)
    return FakeEngine


def _tts(**kw):
    return infer.SmallTTS(state_dicts=({"a": 1}, {"b": 2}), **kw)


def test_synthesize_matches_the_reference_contract(fake):
    tts = _tts(seed=5)
    ref = np.random.default_rng(0).standard_normal((7, 64)).astype(np.float32)
    audio = tts.synthesize(ref, [11, 12, 13], 2.0)  # infer/onnx.py:84: int(2.0 * 24000 / 3200) = 15 frames
    assert audio.shape == (1, 15 * 3200) and audio.dtype == np.float32 and float(audio[0, 0]) == 11.0
    assert infer.frames_for(0.01) == 1 and infer.frames_for(10.0) == 75
    c = fake.instances[0].calls[-1]
    assert c == dict(B=1, T=15, frames=[15], seed=5, steps=4, noise=False, ref_len=[7], ph_len=[3])
    tts.synthesize(ref, [11], 2.0)
    assert fake.instances[0].calls[-1]["seed"] == 6  # a fresh Philox stream per call
    tts.synthesize(ref, [11], 2.0, noise=np.zeros((4, 15, 64), np.float32))  # (steps, T, 64) accepted
    assert fake.instances[0].calls[-1]["noise"]
    assert infer.estimate_duration("x" * 23) == 2.0 and infer.estimate_duration("") == 0.5


def test_ragged_batch_is_padded_once_and_trimmed_per_utterance(fake):
    tts = _tts(num_steps=2)
    rng = np.random.default_rng(1)
    refs = [rng.standard_normal((r, 64)).astype(np.float32) for r in (3, 9, 5)]
    out = tts.synthesize_batch(refs, [[5, 6], [7], [8, 9, 10, 11]], [1.0, 0.5, 2.0])
    assert [a.shape for a in out] == [(1, 7 * 3200), (1, 3 * 3200), (1, 15 * 3200)]
    assert [float(a[0, -1]) for a in out] == [5.0, 7.0, 8.0]
    c = fake.instances[0].calls[-1]
    assert c["B"] == 3 and c["T"] == 15 and c["ref_len"] == [3, 9, 5] and c["ph_len"] == [2, 1, 4] and c["steps"] == 2
    with pytest.raises(ValueError):
        tts.synthesize_batch(refs, [[1]], [1.0])
    with pytest.raises(ValueError):
        tts.synthesize_batch([np.zeros((3, 32), np.float32)], [[1]], [1.0])


def test_forward_concatenates_transcription_and_text_tokens(fake):
    import torch

    tts = _tts()
    conds = [torch.zeros(4, 64), torch.zeros(6, 64)]
    out = tts(conds, [[1, 2], [3]], [[9], [8, 7]], duration_sec=1.0)  # infer/onnx.py:131-157, __call__ = forward
    assert len(out) == 2 and all(isinstance(a, torch.Tensor) and tuple(a.shape) == (1, 7 * 3200) for a in out)
    assert fake.instances[0].calls[-1]["ph_len"] == [3, 3]
    assert tts.forward([], [], []) == []
    with pytest.raises(RuntimeError, match="token ids"):
        tts(conds[:1], ["hello"], ["world"])  # strings need the reference's espeak front-end


def test_devices_split_uses_every_gpu_and_keeps_the_order(fake):
    tts = _tts(devices=[0, 2, 5])
    assert [e.device for e in fake.instances] == [0, 2, 5] and len(tts._replicas) == 2
    rng = np.random.default_rng(2)
    n = 9
    refs = [rng.standard_normal((4, 64)).astype(np.float32) for _ in range(n)]
    ids = [[i + 1] for i in range(n)]
    durs = [1.0 + (i % 3) for i in range(n)]
    noise = rng.standard_normal((4, n, 22, 64)).astype(np.float32)
    out = tts.synthesize_batch(refs, ids, durs, noise=noise)
    assert [a.shape[1] for a in out] == [infer.frames_for(d) * 3200 for d in durs]
    assert [int(a[0, 0]) % 1000 for a in out] == list(range(1, n + 1))  # input order kept
    assert {int(a[0, 0]) // 1000 for a in out} == {0, 2, 5}  # every GPU served a shard
    assert all(c["noise"] for e in fake.instances for c in e.calls)
    assert sum(c["B"] for e in fake.instances for c in e.calls) == n
    assert len(tts._replicas) == 2  # restored after the call
    one = tts.synthesize_batch(refs[:1], ids[:1], durs[:1])  # a single utterance stays on the primary
    assert int(one[0][0, 0]) == 1
    with pytest.raises(ValueError):
        _tts(devices=[1, 1])
    with pytest.raises(ValueError):
        _tts(devices=[])


def test_shape_buckets_round_the_padded_shape_up(fake):
    """Fewer distinct (R, P, T) -> the engine re-uses its per-shape plans; lengths still mask the padding and every
    utterance keeps its own number of samples."""
    tts = _tts(shape_buckets=(8, 16, 5))
    rng = np.random.default_rng(3)
    refs = [rng.standard_normal((r, 64)).astype(np.float32) for r in (3, 9)]
    out = tts.synthesize_batch(refs, [[5, 6], [7] * 17], [1.0, 2.2])  # frames 7 and 16
    c = fake.instances[0].calls[-1]
    assert c["T"] == 20 and c["frames"] == [7, 16] and c["ref_len"] == [3, 9] and c["ph_len"] == [2, 17]
    assert [a.shape for a in out] == [(1, 7 * 3200), (1, 16 * 3200)]
    assert fake.instances[0].last_shape == (16, 32)  # R 9 -> 16, P 17 -> 32
    noise = np.zeros((4, 2, 16, 64), np.float32)
    tts.synthesize_batch(refs, [[5, 6], [7] * 17], [1.0, 2.2], noise=noise)  # supplied noise pins T
    assert fake.instances[0].calls[-1]["T"] == 16
    with pytest.raises(ValueError):
        _tts(shape_buckets=(8, 16))

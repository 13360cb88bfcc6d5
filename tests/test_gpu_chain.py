"""Chained DiT GEMM kernel (csrc/dit_chain.cu) against torch on a real B200, through the C ABI's test hooks.

The kernel folds AdaLayerNormZero, per-head RMSNorm + RoPE, SwiGLU, the gated residuals and the row mask of a DiT block
(dit.py:12-25,95-135,176-212) into GEMM epilogues and replaces kernel boundaries by in-kernel ready counters.  Every
phase is checked on its own, then the four-phase chain (to_out -> w1|w3 -> w2 -> q|k|v|gate of the next block) against
the same phases run one launch at a time and against the block's textbook definition in fp32."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

D, H, HD, HDP, FF, NBLK = 960, 8, 120, 128, 2400, 12
NQ, N13, PARTS = 3 * H * HDP + D, 2 * FF, 30
MOD_LD = NBLK * 6 * D + 2 * D
QKVG, OUT, W13, W2, VEL, ATTN = range(6)


@pytest.fixture(scope="module")
def eng():
    from smalltts_b200.engine import Engine

    e = Engine(0)
    yield e
    e.close()


def _ptr(t):
    if t is None:
        return None
    torch.cuda.current_stream(t.device).synchronize()
    return C.c_void_p(t.data_ptr())


def _bf(x):
    return x.to(torch.bfloat16)


class Case:
    """Random weights of TWO blocks (blk 1 and 2; the rest of the stacked arrays stays zero), one timestep's adaLN
    table and activations for M rows of utterances of T rows."""

    def __init__(self, M, T, frames, seed=0, R=0, P=0, ref_len=None, ph_len=None):
        g = torch.Generator(device="cuda").manual_seed(seed)
        r = lambda *s, scale=1.0: torch.randn(*s, device="cuda", generator=g) * scale  # noqa: E731
        self.M, self.T = M, T
        self.frames = torch.tensor(frames, dtype=torch.int32, device="cuda")
        self.wqkvg = torch.zeros(NBLK, NQ, D, device="cuda", dtype=torch.bfloat16)
        self.bqkvg = torch.zeros(NBLK, NQ, device="cuda")
        self.wo = torch.zeros(NBLK, D, H * HDP, device="cuda", dtype=torch.bfloat16)
        self.w13 = torch.zeros(NBLK, N13, D, device="cuda", dtype=torch.bfloat16)
        self.b13 = torch.zeros(NBLK, N13, device="cuda")
        self.w2 = torch.zeros(NBLK, D, FF, device="cuda", dtype=torch.bfloat16)
        self.b2 = torch.zeros(NBLK, D, device="cuda")
        self.qn = torch.zeros(NBLK, H, HD, device="cuda")
        self.kn = torch.zeros(NBLK, H, HD, device="cuda")
        head_rows = (torch.arange(H * HD, device="cuda") // HD) * HDP + torch.arange(H * HD, device="cuda") % HD
        self.dense = {}  # unpadded fp32 copies (bf16-rounded values) for the textbook reference
        for blk in (1, 2):
            dq = {}
            for j, name in enumerate(("q", "k", "v")):
                w, b = _bf(r(D, D, scale=D ** -0.5)), r(D, scale=0.3)
                self.wqkvg[blk, j * H * HDP + head_rows] = w
                self.bqkvg[blk, j * H * HDP + head_rows] = b
                dq[name] = (w.float(), b)
            wg = _bf(r(D, D, scale=D ** -0.5))
            self.wqkvg[blk, 3 * H * HDP:] = wg
            dq["g"] = (wg.float(), None)
            wo = _bf(r(D, D, scale=D ** -0.5))
            cols = head_rows  # K of to_out is head-padded the same way
            self.wo[blk][:, cols] = wo
            dq["o"] = wo.float()
            w1, w3 = _bf(r(FF, D, scale=D ** -0.5)), _bf(r(FF, D, scale=D ** -0.5))
            b1, b3 = r(FF, scale=0.3), r(FF, scale=0.3)
            il = (torch.arange(FF, device="cuda") // 16) * 32 + torch.arange(FF, device="cuda") % 16
            self.w13[blk, il], self.w13[blk, il + 16] = w1, w3
            self.b13[blk, il], self.b13[blk, il + 16] = b1, b3
            dq["w1"], dq["w3"], dq["b1"], dq["b3"] = w1.float(), w3.float(), b1, b3
            w2 = _bf(r(D, FF, scale=FF ** -0.5))
            self.w2[blk] = w2
            self.b2[blk] = r(D, scale=0.3)
            dq["w2"] = w2.float()
            self.qn[blk], self.kn[blk] = 1 + r(H, HD, scale=0.2), 1 + r(H, HD, scale=0.2)
            self.dense[blk] = dq
        self.wvel, self.bvel = _bf(r(64, D, scale=D ** -0.5)), r(64, scale=0.3)
        pos = torch.arange(4096, device="cuda", dtype=torch.float32)[:, None]
        inv = 1.0 / (10000.0 ** (torch.arange(0, 64, 2, device="cuda", dtype=torch.float32) / 64))
        self.cos_t, self.sin_t = torch.cos(pos * inv).contiguous(), torch.sin(pos * inv).contiguous()
        mod = r(MOD_LD, scale=0.3)
        for blk in range(NBLK):  # gates are stored tanh'ed (engine.cu compute_mod)
            for ch in (2, 5):
                mod[blk * 6 * D + ch * D: blk * 6 * D + (ch + 1) * D].tanh_()
        self.mod = mod
        self.x = (r(M, D, scale=1.5) + 0.4 * r(M, 1)).contiguous()
        self.ob = torch.zeros(M, H * HDP, device="cuda", dtype=torch.bfloat16)
        self.ob.view(M, H, HDP)[:, :, :HD] = _bf(r(M, H, HD))
        self.xb = torch.zeros(M, D, device="cuda", dtype=torch.bfloat16)
        self.stats = torch.zeros(M, PARTS, 2, device="cuda")
        self.qkv = torch.zeros(2, 3, M, H * HDP, device="cuda", dtype=torch.bfloat16)  # [block parity] when qkv_db
        self.qkv_db = 0
        # cross-attention caches of the conditions (only read by ATTN phases)
        self.B, self.R, self.P = M // T, R, P
        self.kv_ref = self.kv_text = self.ref_len = self.ph_len = None
        if R:
            self.kv_ref = torch.zeros(NBLK, 2, self.B, R, H, HDP, device="cuda", dtype=torch.bfloat16)
            self.kv_text = torch.zeros(NBLK, 2, self.B, P, H, HDP, device="cuda", dtype=torch.bfloat16)
            self.kv_ref[..., :HD] = _bf(r(NBLK, 2, self.B, R, H, HD))
            self.kv_text[..., :HD] = _bf(r(NBLK, 2, self.B, P, H, HD))
            self.ref_len = torch.tensor(ref_len, dtype=torch.int32, device="cuda")
            self.ph_len = torch.tensor(ph_len, dtype=torch.int32, device="cuda")
        self.gate = torch.zeros(M, D, device="cuda")
        self.hb = torch.zeros(M, FF, device="cuda", dtype=torch.bfloat16)
        self.vel = torch.zeros(M, 64, device="cuda")
        self.fold = None

    # ---- pieces of the adaLN table
    def m(self, blk, ch):
        return self.mod[blk * 6 * D + ch * D: blk * 6 * D + (ch + 1) * D]

    def final(self, ch):
        return self.mod[NBLK * 6 * D + ch * D: NBLK * 6 * D + (ch + 1) * D]

    def args(self, phases):
        from smalltts_b200 import _cabi

        a = _cabi.ChainArgs()
        for n in ("wqkvg", "wo", "w13", "w2", "wvel", "bqkvg", "b13", "b2", "bvel", "qn", "kn", "cos_t", "sin_t", "x", "xb",
                  "stats", "qkv", "gate", "ob", "hb", "vel", "frames", "mod", "fold", "kv_ref", "kv_text", "ref_len", "ph_len"):
            t = getattr(self, n)
            setattr(a, n, None if t is None else _ptr(t))
        n_ready = (len(phases) * ((self.M + 127) // 128) + 31) // 32 * 32 + 64  # dit_chain.cuh chain_ready_ints
        self.ready = torch.zeros(n_ready, dtype=torch.int32, device="cuda")
        a.ready = _ptr(self.ready)
        a.M, a.T, a.n_phases = self.M, self.T, len(phases)
        a.B, a.R, a.P, a.qkv_db = self.B, self.R, self.P, self.qkv_db
        for i, (k, b) in enumerate(phases):
            a.kind[i], a.blk[i] = k, b
        return a

    def run(self, eng, phases):
        from smalltts_b200 import _cabi

        a = self.args(phases)
        _cabi.check(_cabi.lib().stts_test_chain(eng._h, C.byref(a)), eng._h)
        torch.cuda.synchronize()

    def make_fold(self, eng):
        from smalltts_b200 import _cabi

        n = int(_cabi.lib().stts_test_chain_fold_floats())
        self.fold = torch.zeros(n, device="cuda")
        a = self.args([(QKVG, 1)])
        _cabi.check(_cabi.lib().stts_test_chain_fold(eng._h, C.byref(a), _ptr(self.fold)), eng._h)
        torch.cuda.synchronize()

    def stats_cast(self, eng, scale):
        from smalltts_b200 import _cabi

        _cabi.check(_cabi.lib().stts_test_chain_stats_cast(eng._h, _ptr(self.x), self.M, _ptr(scale.contiguous()), _ptr(self.xb),
                                                           _ptr(self.stats)), eng._h)
        torch.cuda.synchronize()


def _ln(x):
    return torch.nn.functional.layer_norm(x, (D,), eps=1e-6)


def _rope(v, T):  # v [M, H, HD]: interleaved pairs of the first 64 dims, position = row % T (dit.py:152-173)
    M = v.shape[0]
    pos = (torch.arange(M, device=v.device) % T).float()[:, None]
    inv = 1.0 / (10000.0 ** (torch.arange(0, 64, 2, device=v.device, dtype=torch.float32) / 64))
    ang = pos * inv  # [M, 32]
    c, s = torch.cos(ang)[:, None, :], torch.sin(ang)[:, None, :]
    x0, x1 = v[..., 0:64:2], v[..., 1:64:2]
    out = v.clone()
    out[..., 0:64:2] = x0 * c - x1 * s
    out[..., 1:64:2] = x1 * c + x0 * s
    return out


def _rms_heads(v, w):  # v [M, H, HD], w [H, HD]  (dit.py:52-53, eps 1e-6)
    return v * torch.rsqrt(v.pow(2).mean(-1, keepdim=True) + 1e-6) * w


def ref_qkvg(c, blk, x):
    """Textbook q|k|v|gate of block blk from the fp32 residual x: -> q, k, v [M, H, HD] and gate [M, D]."""
    d = c.dense[blk]
    n = _ln(x) * (1 + c.m(blk, 1)) + c.m(blk, 0)
    q = (n @ d["q"][0].t() + d["q"][1]).view(-1, H, HD)
    k = (n @ d["k"][0].t() + d["k"][1]).view(-1, H, HD)
    v = (n @ d["v"][0].t() + d["v"][1]).view(-1, H, HD)
    g = n @ d["g"][0].t()
    return _rope(_rms_heads(q, c.qn[blk]), c.T), _rope(_rms_heads(k, c.kn[blk]), c.T), v, g


def ref_out(c, blk, x, ob):
    a = ob.float().view(-1, H, HDP)[:, :, :HD].reshape(-1, D) @ c.dense[blk]["o"].t()
    rows = torch.arange(c.M, device="cuda")
    live = (rows % c.T) < c.frames[rows // c.T]
    a = a * live[:, None]
    return x + c.m(blk, 2) * a


def ref_mlp(c, blk, x):
    d = c.dense[blk]
    n = _ln(x) * (1 + c.m(blk, 4)) + c.m(blk, 3)
    h = torch.nn.functional.silu(n @ d["w1"].t() + d["b1"]) * (n @ d["w3"].t() + d["b3"])
    return x + c.m(blk, 5) * (h @ d["w2"].t() + c.b2[blk]), h


def rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))


def check_stats(c, x):
    parts = x.view(c.M, PARTS, 32)
    want = torch.stack([parts.sum(-1), parts.pow(2).sum(-1)], dim=-1)
    assert rel(c.stats, want) < 1e-5


@pytest.mark.parametrize("M,T,frames", [(600, 75, [75] * 8), (210, 70, [70, 33, 1])])
def test_fold_table_stats_cast_and_qkvg_phase(eng, M, T, frames):
    c = Case(M, T, frames, seed=1)
    c.make_fold(eng)
    # fold vectors of block 1: cs = W (1 + scale), b' = b + W shift   (bf16-rounded W, as packed)
    W = c.wqkvg[1].float()
    assert rel(c.fold[1 * NQ: 2 * NQ], W @ (1 + c.m(1, 1))) < 1e-5
    assert rel(c.fold[NBLK * NQ + 1 * NQ: NBLK * NQ + 2 * NQ], c.bqkvg[1] + W @ c.m(1, 0)) < 1e-5
    W = c.w13[2].float()
    off = 2 * NBLK * NQ
    assert rel(c.fold[off + 2 * N13: off + 3 * N13], W @ (1 + c.m(2, 4))) < 1e-5
    assert rel(c.fold[off + NBLK * N13 + 2 * N13: off + NBLK * N13 + 3 * N13], c.b13[2] + W @ c.m(2, 3)) < 1e-5
    off = 2 * NBLK * NQ + 2 * NBLK * N13
    assert rel(c.fold[off: off + 64], c.wvel.float() @ (1 + c.final(0))) < 1e-5
    assert rel(c.fold[off + 64: off + 128], c.bvel + c.wvel.float() @ c.final(1)) < 1e-5

    c.stats_cast(eng, c.m(1, 1))
    check_stats(c, c.x)
    assert rel(c.xb, c.x * (1 + c.m(1, 1))) < 4e-3  # bf16 rounding

    c.run(eng, [(QKVG, 1)])
    q, k, v, g = ref_qkvg(c, 1, c.x)
    got = c.qkv[0].float().view(3, M, H, HDP)
    assert float(got[..., HD:].abs().max()) == 0.0  # head padding stays zero
    for i, (name, want) in enumerate((("q", q), ("k", k), ("v", v))):
        err = rel(got[i, :, :, :HD], want)
        print("qkvg", name, err)
        assert err < 1e-2, (name, err)
    err = rel(c.gate, g)
    print("gate", err)
    assert err < 1e-2


@pytest.mark.parametrize("M,T,frames", [(600, 75, [75, 75, 60, 75, 1, 75, 75, 75]), (130, 65, [65, 20])])
def test_each_phase_alone(eng, M, T, frames):
    c = Case(M, T, frames, seed=2)
    c.make_fold(eng)
    x0 = c.x.clone()
    # to_out of block 1: x1 = x0 + tanh(gate_msa) * mask(Wo o); operand for w1|w3 is bf16(x1 (1 + scale_mlp))
    c.run(eng, [(OUT, 1)])
    x1 = ref_out(c, 1, x0, c.ob)
    assert rel(c.x, x1) < 2e-3
    check_stats(c, c.x)
    assert rel(c.xb, c.x * (1 + c.m(1, 4))) < 4e-3
    # w1|w3 of block 1 on that x
    x1g = c.x.clone()
    c.run(eng, [(W13, 1)])
    x2, h = ref_mlp(c, 1, x1g)
    err = rel(c.hb, h)
    print("hidden", err)
    assert err < 1.5e-2
    # w2 of block 1: x2 = x1 + tanh(gate_mlp) (W2 h + b2); operand for block 2's q|k|v|gate uses scale_msa of block 2
    hb = c.hb.float()
    c.run(eng, [(W2, 1)])
    want = x1g + c.m(1, 5) * (hb @ c.dense[1]["w2"].t() + c.b2[1])
    assert rel(c.x, want) < 2e-3
    check_stats(c, c.x)
    assert rel(c.xb, c.x * (1 + c.m(2, 1))) < 4e-3
    # velocity head behind the final adaLN (needs xb scaled with the FINAL scale: re-cast)
    c.stats_cast(eng, c.final(0))
    c.run(eng, [(VEL, 0)])
    vel = (_ln(c.x) * (1 + c.final(0)) + c.final(1)) @ c.wvel.float().t() + c.bvel
    err = rel(c.vel, vel)
    print("velocity", err)
    assert err < 1e-2


@pytest.mark.parametrize("M,T,frames", [(600, 75, [75] * 8), (600, 75, [75, 40, 75, 75, 9, 75, 75, 75]), (1500, 150, [150] * 10),
                                        (70, 70, [64])])
def test_four_phase_chain_equals_single_launches_and_the_textbook_block(eng, M, T, frames):
    """to_out(1) -> w1|w3(1) -> w2(1) -> q|k|v|gate(2) in ONE launch (row-block counters instead of kernel boundaries)
    must give bit-identical results to the same four phases launched one by one, and match the fp32 definition."""
    c = Case(M, T, frames, seed=3)
    c.make_fold(eng)
    x0 = c.x.clone()
    for ph in ((OUT, 1), (W13, 1), (W2, 1), (QKVG, 2)):
        c.run(eng, [ph])
    single = [t.clone() for t in (c.x, c.xb, c.stats, c.hb, c.qkv, c.gate)]
    for rep in range(3):  # repeat: scheduling differs from run to run, results must not
        c.x.copy_(x0)
        for t in (c.xb, c.stats, c.hb, c.qkv, c.gate):
            t.zero_()
        c.run(eng, [(OUT, 1), (W13, 1), (W2, 1), (QKVG, 2)])
        for name, got, want in zip(("x", "xb", "stats", "hb", "qkv", "gate"), (c.x, c.xb, c.stats, c.hb, c.qkv, c.gate), single):
            assert torch.equal(got, want), (rep, name, float((got.float() - want.float()).abs().max()))
    x1 = ref_out(c, 1, x0, c.ob)
    x2, _ = ref_mlp(c, 1, x1)
    assert rel(c.x, x2) < 1e-2
    q, k, v, g = ref_qkvg(c, 2, x2)
    got = c.qkv[0].float().view(3, M, H, HDP)
    for i, want in enumerate((q, k, v)):
        assert rel(got[i, :, :, :HD], want) < 2e-2, i
    assert rel(c.gate, g) < 2e-2


def test_last_chain_ends_in_the_velocity_head(eng):
    c = Case(600, 75, [75] * 8, seed=4)
    c.make_fold(eng)
    x0 = c.x.clone()
    # the last block's w2 epilogue scales its bf16 copy with the FINAL adaLN scale
    for blk_arrays in (c.wo, c.w13, c.w2, c.b13, c.b2):
        blk_arrays[NBLK - 1] = blk_arrays[1]
    c.dense[NBLK - 1] = c.dense[1]
    c.mod[(NBLK - 1) * 6 * D: NBLK * 6 * D] = c.mod[6 * D: 12 * D]
    c.make_fold(eng)
    c.run(eng, [(OUT, NBLK - 1), (W13, NBLK - 1), (W2, NBLK - 1), (VEL, 0)])
    x1 = ref_out(c, NBLK - 1, x0, c.ob)
    x2, _ = ref_mlp(c, NBLK - 1, x1)
    vel = (_ln(x2) * (1 + c.final(0)) + c.final(1)) @ c.wvel.float().t() + c.bvel
    assert rel(c.x, x2) < 1e-2
    err = rel(c.vel, vel)
    print("chain velocity", err)
    assert err < 2e-2


# ---------------------------------------------------------------- attention as a phase of the chain (chain_attn.cuh)
def ref_attention(c, blk, qkv, gate):
    """Joint self | ref | text attention of block blk (dit.py:110-119,131-135) from bf16 q|k|v [3, M, H*HDP] and the
    pre-sigmoid gate [M, D]: -> [M, H*HDP] fp32 (head padding zero).  Keys beyond the valid lengths are masked; query rows
    beyond frames[b] are computed like any other (to_out masks them)."""
    M, T, B = c.M, c.T, c.B
    q, k, v = (qkv[i].float().view(B, T, H, HDP) for i in range(3))
    out = torch.zeros(B, T, H, HDP, device="cuda")
    for b in range(B):
        n0, n1, n2 = int(c.frames[b]), int(c.ref_len[b]), int(c.ph_len[b])
        ks = torch.cat([k[b, :n0], c.kv_ref[blk, 0, b, :n1].float(), c.kv_text[blk, 0, b, :n2].float()])  # [N, H, HDP]
        vs = torch.cat([v[b, :n0], c.kv_ref[blk, 1, b, :n1].float(), c.kv_text[blk, 1, b, :n2].float()])
        sc = torch.einsum("qhd,khd->hqk", q[b], ks) * HD ** -0.5
        out[b] = torch.einsum("hqk,khd->qhd", sc.softmax(-1), vs)
    out = out.view(M, H, HDP)
    out[:, :, :HD] *= torch.sigmoid(gate.view(M, H, HD))
    return out.view(M, H * HDP)


ATTN_CASES = [
    # M, T, frames, R, P, ref_len, ph_len
    (640, 80, [80] * 8, 75, 60, [75, 60, 75, 30, 75, 75, 1, 75], [60, 55, 41, 60, 13, 60, 60, 60]),
    (210, 70, [70, 33, 1], 20, 17, [20, 1, 7], [17, 16, 1]),
    (450, 150, [150, 97, 131], 40, 30, [40, 17, 33], [30, 3, 16]),  # two query tiles per utterance; 160 + 48 + 32 keys
]


@pytest.mark.parametrize("M,T,frames,R,P,ref_len,ph_len", ATTN_CASES)
@pytest.mark.parametrize("blk", [1, 2])
def test_attention_phase_alone(eng, M, T, frames, R, P, ref_len, ph_len, blk):
    c = Case(M, T, frames, seed=5, R=R, P=P, ref_len=ref_len, ph_len=ph_len)
    c.make_fold(eng)
    c.qkv_db = 1
    g = torch.Generator(device="cuda").manual_seed(50 + blk)
    par = blk & 1
    c.qkv[par].view(3, M, H, HDP)[..., :HD] = _bf(torch.randn(3, M, H, HD, device="cuda", generator=g))
    # the OTHER half holds NaNs: an item must never touch it
    c.qkv[1 - par].fill_(float("nan"))
    c.gate.copy_(torch.randn(M, D, device="cuda", generator=g))
    c.ob.fill_(float("nan"))
    c.run(eng, [(ATTN, blk)])
    want = ref_attention(c, blk, c.qkv[par], c.gate)
    assert torch.isfinite(c.ob.float()).all()
    assert float(c.ob.float().view(M, H, HDP)[..., HD:].abs().max()) == 0.0
    err = rel(c.ob, want)
    print("attention phase", err)
    assert err < 1e-2
    counts = c.ready[: (M + 127) // 128].tolist()  # items per row block: 8 heads x query tiles touching the block
    want_counts = [0] * ((M + 127) // 128)
    for b in range(M // T):
        for q0 in range(0, T, 128):
            r0, r1 = b * T + q0, b * T + min(q0 + 128, T) - 1
            for m in range(r0 // 128, r1 // 128 + 1):
                want_counts[m] += H
    assert counts == want_counts


@pytest.mark.parametrize("M,T,frames,R,P,ref_len,ph_len", ATTN_CASES)
def test_whole_block_in_one_launch_equals_phase_by_phase(eng, M, T, frames, R, P, ref_len, ph_len):
    """q|k|v|gate(1) -> attention(1) -> to_out(1) -> w1|w3(1) -> w2(1) -> q|k|v|gate(2) -> attention(2) in ONE launch is
    bit-identical to the same phases launched one at a time, and matches the textbook block."""
    c = Case(M, T, frames, seed=6, R=R, P=P, ref_len=ref_len, ph_len=ph_len)
    c.make_fold(eng)
    c.qkv_db = 1
    c.stats_cast(eng, c.m(1, 1))
    x0 = c.x.clone()
    xb0, stats0 = c.xb.clone(), c.stats.clone()
    phases = [(QKVG, 1), (ATTN, 1), (OUT, 1), (W13, 1), (W2, 1), (QKVG, 2), (ATTN, 2)]
    bufs = ("x", "xb", "stats", "hb", "qkv", "gate", "ob")

    def reset():
        c.x.copy_(x0), c.xb.copy_(xb0), c.stats.copy_(stats0)
        for n in ("hb", "qkv", "gate", "ob"):
            getattr(c, n).zero_()

    reset()
    for i, ph in enumerate(phases):
        c.run(eng, [ph])
        if i == 1:
            ob1 = c.ob.clone()  # attention of block 1, consumed by to_out
    single = [getattr(c, n).clone() for n in bufs]
    for rep in range(3):
        reset()
        c.run(eng, phases)
        for n, want in zip(bufs, single):
            got = getattr(c, n)
            assert torch.equal(got, want), (rep, n, float((got.float() - want.float()).abs().max()))
    # textbook: block 1 from x0, then q|k|v|gate and attention of block 2
    q, k, v, g = ref_qkvg(c, 1, x0)
    assert rel(c.qkv[1].float().view(3, M, H, HDP)[0, :, :, :HD], q) < 1e-2
    a1 = ref_attention(c, 1, c.qkv[1], g)
    assert rel(ob1, a1) < 2e-2
    x1 = ref_out(c, 1, x0, ob1)
    x2, _ = ref_mlp(c, 1, x1)
    assert rel(c.x, x2) < 1e-2
    q, k, v, g = ref_qkvg(c, 2, x2)
    assert rel(c.gate, g) < 2e-2
    a2 = ref_attention(c, 2, c.qkv[0], c.gate)
    err = rel(c.ob, a2)
    print("block in one launch: attention of the next block", err)
    assert err < 1e-2

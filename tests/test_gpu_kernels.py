"""Kernel-level parity on a real B200, through the C ABI's test hooks.  Reference values are computed by torch
on the same bf16-rounded operands in fp32, so tolerances only cover accumulation order and the bf16 rounding of
outputs where the kernel emits bf16."""
import ctypes as C
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

ACT = dict(none=0, gelu=1, mish=2, sigmoid=3, swiglu=4, silu=5)


@pytest.fixture(scope="module")
def eng():
    from smalltts_b200.engine import Engine

    e = Engine(0)
    yield e
    e.close()


def _p(t):
    """Raw device pointer for a C-ABI test hook.  The engine runs on its own non-blocking stream, so everything torch
    has queued for this tensor (randn, zeros, casts) must have finished before the hook launches: synchronise here."""
    if t is None:
        return None
    torch.cuda.current_stream(t.device).synchronize()
    return C.c_void_p(t.data_ptr())


def run_gemm(eng, a, w, *, B, T, N, K, bn, taps=1, shift0=0, step=1, groups=1, a_koff=0, w_grows=0, out_gcols=0,
             bias=None, act="none", row_len=None, rpb=0, mask_bf16_only=0, colscale=None, rowgate=None, ld_gate=0,
             residual=None, ld_out=None, out_cols=None, want_bf16=False):
    from smalltts_b200 import _cabi

    out_cols = out_cols or (N * groups if groups > 1 else N)
    ld_out = ld_out or out_cols
    out32 = torch.zeros(B * T, ld_out, device="cuda", dtype=torch.float32)
    out16 = torch.zeros(B * T, ld_out, device="cuda", dtype=torch.bfloat16) if want_bf16 else None
    rc = _cabi.lib().stts_test_gemm(
        eng._h, bn, _p(a), B, T, a.shape[-1], a.stride(-2), _p(w), w.shape[0], w.stride(0), N, K, taps, shift0, step,
        groups, a_koff, w_grows, out_gcols, _p(bias), ACT[act], _p(row_len), rpb, mask_bf16_only, _p(colscale), _p(rowgate),
        ld_gate, _p(residual), residual.stride(0) if residual is not None else 0, _p(out32), _p(out16), ld_out)
    _cabi.check(rc, eng._h)
    torch.cuda.synchronize()
    return out32, out16


def _rand_bf16(*shape, scale=1.0):
    return (torch.randn(*shape, device="cuda") * scale).to(torch.bfloat16)


def _assert_close(got, want, tol=2e-3):
    err = (got.float() - want.float()).abs().max().item()
    ref = want.float().abs().max().item() + 1e-6
    assert err <= tol * ref, f"max abs err {err:.3e} vs scale {ref:.3e}"


@pytest.mark.parametrize("bn", [32, 64, 128, 256])
@pytest.mark.parametrize("M,N,K", [(600, 960, 960), (75, 64, 960), (1000, 4096, 512), (130, 960, 2400)])
def test_linear(eng, bn, M, N, K):
    torch.manual_seed(0)
    a, w = _rand_bf16(M, K), _rand_bf16(N, K, scale=K ** -0.5)
    bias = torch.randn(N, device="cuda")
    out, _ = run_gemm(eng, a, w, B=1, T=M, N=N, K=K, bn=bn, bias=bias)
    _assert_close(out, a.float() @ w.float().t() + bias)


@pytest.mark.parametrize("M,N,K,bn,splits,gelu2", [(600, 2048, 8192, 128, 3, False), (600, 8192, 2048, 256, 3, True),
                                                      (4800, 1024, 4096, 256, 3, False), (600, 8192, 4096, 256, 2, False),
                                                      (130, 256, 1024, 128, 4, False)])
def test_linear_split_k(eng, M, N, K, bn, splits, gelu2):
    """Split-K (gemm.cuh GemmShape::splits): every tile is computed by `splits` work items over parts of the reduction;
    the last part to arrive adds the parked fp32 partials in part order and runs the epilogue.  Same result as the
    whole-tile GEMM up to fp32 summation order, bit-identical from run to run, counters left clean (the hook launches
    twice on one counter buffer)."""
    from smalltts_b200 import _cabi

    torch.manual_seed(7)
    a, w = _rand_bf16(M, K), _rand_bf16(N, K, scale=K ** -0.5)
    bias = torch.randn(N, device="cuda")
    if gelu2:  # the vocoder's FFN1 epilogue: 2 * gelu(x) written as fp16
        outs = []
        for _ in range(2):
            o16 = torch.zeros(M, N, device="cuda", dtype=torch.float16)
            rc = _cabi.lib().stts_test_gemm_split(eng._h, bn, splits, _p(a), M, K, _p(w), N, _p(bias), 1, None, None, None, _p(o16))
            _cabi.check(rc, eng._h)
            torch.cuda.synchronize()
            outs.append(o16)
        want = 2 * torch.nn.functional.gelu(a.float() @ w.float().t() + bias)
        _assert_close(outs[0], want, tol=4e-3)
    else:
        cs, res = torch.rand(N, device="cuda") + 0.5, torch.randn(M, N, device="cuda")
        outs = []
        for _ in range(2):
            o32 = torch.zeros(M, N, device="cuda")
            rc = _cabi.lib().stts_test_gemm_split(eng._h, bn, splits, _p(a), M, K, _p(w), N, _p(bias), 0, _p(cs), _p(res), _p(o32), None)
            _cabi.check(rc, eng._h)
            torch.cuda.synchronize()
            outs.append(o32)
        _assert_close(outs[0], (a.float() @ w.float().t() + bias) * cs + res)
    assert torch.equal(outs[0], outs[1])


PAIR = 0x1000  # gemm.cuh kGemmPairFlag: clusters of two CTAs, 256 x bn tiles, cta_group::2 MMAs
# The engine uses the pair variant only with STTS_GEMM_2CTA=1 (measured: wins on long reductions in isolation, no
# end-to-end gain yet); its kernel tests run by default (STTS_TEST_PAIR=0 skips them).
pair_only = pytest.mark.skipif(os.environ.get("STTS_TEST_PAIR") == "0", reason="CTA-pair GEMM tests disabled: STTS_TEST_PAIR=0")


@pair_only

@pytest.mark.parametrize("bn", [128, 256])
@pytest.mark.parametrize("M,N,K", [(600, 960, 960), (75, 256, 64), (20000, 1024, 256), (1000, 4096, 512), (130, 960, 2400)])
def test_linear_cta_pair(eng, bn, M, N, K):
    """CTA-pair variant of the GEMM (each CTA stages 128 rows of A and bn/2 rows of W; the leader issues M = 256 MMAs):
    ragged M (tiles whose second half is empty), N not a multiple of bn, several tiles per pair."""
    torch.manual_seed(0)
    a, w = _rand_bf16(M, K), _rand_bf16(N, K, scale=K ** -0.5)
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda")
    out, _ = run_gemm(eng, a, w, B=1, T=M, N=N, K=K, bn=bn | PAIR, bias=bias, residual=res)
    _assert_close(out, a.float() @ w.float().t() + bias + res)


@pair_only
def test_conv_taps_and_batches_cta_pair(eng):
    """7-tap causal conv over B batches of T rows (the vocoder stem's shape family) on CTA pairs: 256-row tiles must
    not cross batch boundaries and rows shifted before the start of a batch read zero."""
    torch.manual_seed(4)
    B, T, Cin, Cout = 3, 300, 64, 512
    x = _rand_bf16(B, T, Cin)
    wc = _rand_bf16(Cout, Cin, 7, scale=(7 * Cin) ** -0.5)
    bias = torch.randn(Cout, device="cuda")
    w = wc.permute(0, 2, 1).reshape(Cout, 7 * Cin).contiguous()
    out, _ = run_gemm(eng, x, w, B=B, T=T, N=Cout, K=Cin, bn=256 | PAIR, taps=7, shift0=-6, step=1, bias=bias)
    want = torch.nn.functional.conv1d(torch.nn.functional.pad(x.float().transpose(1, 2), (6, 0)), wc.float(), bias)
    _assert_close(out.view(B, T, Cout), want.transpose(1, 2))


def test_small_k_box_exceeds_extent(eng):
    """Vocoder C=32: the 64-wide TMA box is wider than the tensor; the overhang must read as zero."""
    torch.manual_seed(1)
    M, N, K = 5000, 128, 32
    a, w = _rand_bf16(M, K), _rand_bf16(N, K)
    out, out16 = run_gemm(eng, a, w, B=1, T=M, N=N, K=K, bn=128, act="gelu", want_bf16=True)
    want = torch.nn.functional.gelu(a.float() @ w.float().t())
    _assert_close(out, want)
    _assert_close(out16, want, tol=1e-2)


def test_swiglu_interleaved(eng):
    torch.manual_seed(2)
    M, K, Hd = 300, 960, 2400
    a = _rand_bf16(M, K)
    w1, w3 = _rand_bf16(Hd, K, scale=K ** -0.5), _rand_bf16(Hd, K, scale=K ** -0.5)
    b1, b3 = torch.randn(Hd, device="cuda"), torch.randn(Hd, device="cuda")
    idx = torch.arange(Hd, device="cuda")
    lo, hi = (idx // 16) * 32 + idx % 16, (idx // 16) * 32 + 16 + idx % 16
    w = torch.zeros(2 * Hd, K, device="cuda", dtype=torch.bfloat16)
    b = torch.zeros(2 * Hd, device="cuda")
    w[lo], w[hi], b[lo], b[hi] = w1, w3, b1, b3
    _, out16 = run_gemm(eng, a, w, B=1, T=M, N=2 * Hd, K=K, bn=128, bias=b, act="swiglu", want_bf16=True, ld_out=Hd,
                        out_cols=Hd)
    want = torch.nn.functional.silu(a.float() @ w1.float().t() + b1) * (a.float() @ w3.float().t() + b3)
    _assert_close(out16, want, tol=1e-2)


def test_epilogue_mask_gate_scale_residual(eng):
    torch.manual_seed(3)
    B, T, N, K = 3, 75, 960, 1024
    a, w = _rand_bf16(B * T, K), _rand_bf16(N, K, scale=K ** -0.5)
    bias, gate = torch.randn(N, device="cuda"), torch.randn(B, N, device="cuda")
    res = torch.randn(B * T, N, device="cuda")
    lens = torch.tensor([75, 40, 1], device="cuda", dtype=torch.int32)
    out, _ = run_gemm(eng, a.view(B, T, K), w, B=B, T=T, N=N, K=K, bn=64, bias=bias, row_len=lens, rowgate=gate,
                      ld_gate=N, residual=res)
    acc = (a.float() @ w.float().t() + bias).view(B, T, N)
    mask = (torch.arange(T, device="cuda")[None, :] < lens[:, None])[..., None]
    want = res.view(B, T, N) + gate[:, None, :] * (acc * mask)
    _assert_close(out.view(B, T, N), want)
    # flattened rows with rows_per_batch (how the engine's linear layers run), residual updated in place
    out2, _ = run_gemm(eng, a, w, B=1, T=B * T, N=N, K=K, bn=32, bias=bias, row_len=lens, rpb=T, rowgate=gate,
                       ld_gate=N, residual=res)
    _assert_close(out2.view(B, T, N), want)


def test_causal_conv_taps(eng):
    """7-tap causal conv (vocoder stem, hf:181-216) vs F.conv1d with left padding."""
    torch.manual_seed(4)
    B, T, Cin, Cout = 2, 200, 64, 256
    x = _rand_bf16(B, T, Cin)
    wc = _rand_bf16(Cout, Cin, 7, scale=(7 * Cin) ** -0.5)
    bias = torch.randn(Cout, device="cuda")
    w = wc.permute(0, 2, 1).reshape(Cout, 7 * Cin).contiguous()  # [o, tap*Cin + c]
    out, _ = run_gemm(eng, x, w, B=B, T=T, N=Cout, K=Cin, bn=128, taps=7, shift0=-6, step=1, bias=bias)
    want = torch.nn.functional.conv1d(torch.nn.functional.pad(x.float().transpose(1, 2), (6, 0)), wc.float(), bias)
    _assert_close(out.view(B, T, Cout), want.transpose(1, 2))


@pytest.mark.parametrize("r,cin,cout", [(8, 128, 64), (2, 64, 32), (5, 256, 128)])
def test_conv_transpose_as_two_taps(eng, r, cin, cout):
    """CausalConvTranspose1d(k=2r, stride=r) incl. the trim of the last r samples (hf:219-260)."""
    torch.manual_seed(5)
    B, T = 2, 150
    x = _rand_bf16(B, T, cin)
    wt = _rand_bf16(cin, cout, 2 * r, scale=(2 * cin) ** -0.5)
    bias = torch.randn(cout, device="cuda")
    # dst[j*cout + o, tap*cin + c] = w[c, o, j + tap*r]
    w = wt.view(cin, cout, 2, r).permute(3, 1, 2, 0).reshape(r * cout, 2 * cin).contiguous()
    out, _ = run_gemm(eng, x, w, B=B, T=T, N=r * cout, K=cin, bn=128, taps=2, shift0=0, step=-1, bias=bias.repeat(r))
    want = torch.nn.functional.conv_transpose1d(x.float().transpose(1, 2), wt.float(), bias, stride=r)[..., : T * r]
    _assert_close(out.view(B, T * r, cout), want.transpose(1, 2))


def test_grouped_conv_k31(eng):
    """ConvPositionEmbedding convs: Conv1d(960, 960, 31, groups=16, padding=15) + Mish + mask (dit.py:223-236).
    Groups are padded 60 -> 64 channels so that every TMA box start is 16-byte aligned.  conv1 keeps the padded
    layout on its output; conv2 writes the dense [M, 960] stream as 15 tiles of 64 columns, each reading the two
    adjacent padded groups it can touch (K = 128, block-structured weights)."""
    torch.manual_seed(6)
    B, T, Cdim = 2, 75, 960
    lens = torch.tensor([75, 50], device="cuda", dtype=torch.int32)
    keep = (torch.arange(T, device="cuda")[None, :, None] < lens[:, None, None])
    x = _rand_bf16(B, T, Cdim) * keep
    wc = _rand_bf16(Cdim, 60, 31, scale=(60 * 31) ** -0.5)
    bias = torch.randn(Cdim, device="cuda")
    want = torch.nn.functional.mish(
        torch.nn.functional.conv1d(x.float().transpose(1, 2), wc.float(), bias, padding=15, groups=16)).transpose(1, 2)
    want = want * keep
    xp = torch.zeros(B, T, 1024, device="cuda", dtype=torch.bfloat16)
    xp.view(B, T, 16, 64)[..., :60] = x.view(B, T, 16, 60)
    # conv1 style: 16 groups, padded output
    w1 = torch.zeros(16 * 64, 31 * 64, device="cuda", dtype=torch.bfloat16)
    w1.view(16, 64, 31, 64)[:, :60, :, :60] = wc.view(16, 60, 60, 31).permute(0, 1, 3, 2)
    b1 = torch.zeros(1024, device="cuda")
    b1.view(16, 64)[:, :60] = bias.view(16, 60)
    _, o1 = run_gemm(eng, xp, w1, B=B, T=T, N=64, K=64, bn=64, taps=31, shift0=-15, step=1, groups=16, a_koff=64,
                     w_grows=64, out_gcols=64, bias=b1, act="mish", row_len=lens, want_bf16=True, ld_out=1024,
                     out_cols=1024)
    got1 = o1.view(B, T, 16, 64)
    _assert_close(got1[..., :60].reshape(B, T, Cdim), want, tol=1e-2)
    assert got1[..., 60:].abs().max().item() == 0.0
    # conv2 style: dense output tiles
    o = torch.arange(Cdim, device="cuda")
    w2 = torch.zeros(Cdim, 31, 128, device="cuda", dtype=torch.bfloat16)
    kk0 = (o // 60 - o // 64) * 64
    for c in range(60):
        w2[o, :, kk0 + c] = wc[:, c, :]
    w2 = w2.reshape(Cdim, 31 * 128).contiguous()
    res = torch.randn(B * T, Cdim, device="cuda")
    o2, _ = run_gemm(eng, xp, w2, B=B, T=T, N=64, K=128, bn=64, taps=31, shift0=-15, step=1, groups=15, a_koff=64,
                     w_grows=64, out_gcols=64, bias=bias, act="mish", row_len=lens, residual=res, ld_out=Cdim,
                     out_cols=Cdim)
    _assert_close(o2.view(B, T, Cdim), want + res.view(B, T, Cdim), tol=2e-3)


@pytest.mark.parametrize("hd,hd_pad,H", [(120, 128, 8), (64, 64, 8), (128, 128, 4)])
def test_attention(eng, hd, hd_pad, H):
    from smalltts_b200 import _cabi

    torch.manual_seed(7)
    B, Tq, N1 = 3, 75, 140
    def mk(n):
        t = _rand_bf16(B, n, H, hd_pad)
        t[..., hd:] = 0
        return t
    q, k0, v0, k1, v1 = mk(Tq), mk(Tq), mk(Tq), mk(N1), mk(N1)
    len0 = torch.tensor([75, 33, 64], device="cuda", dtype=torch.int32)
    len1 = torch.tensor([140, 1, 77], device="cuda", dtype=torch.int32)
    gate = torch.randn(B * Tq, H * hd, device="cuda")
    out = torch.zeros(B * Tq, H * hd_pad, device="cuda", dtype=torch.bfloat16)
    rc = _cabi.lib().stts_test_attention(eng._h, _p(q), B, Tq, H, hd, hd_pad, _p(k0), _p(v0), _p(len0), Tq, _p(k1),
                                         _p(v1), _p(len1), N1, _p(gate), H * hd, _p(out))
    _cabi.check(rc, eng._h)
    torch.cuda.synchronize()
    kk = torch.cat([k0, k1], 1).float()
    vv = torch.cat([v0, v1], 1).float()
    ok = torch.cat([torch.arange(Tq, device="cuda")[None] < len0[:, None],
                    torch.arange(N1, device="cuda")[None] < len1[:, None]], 1)
    s = torch.einsum("bqhd,bkhd->bhqk", q.float(), kk) / hd ** 0.5
    s = s.masked_fill(~ok[:, None, None, :], float("-inf"))
    o = torch.einsum("bhqk,bkhd->bqhd", s.softmax(-1), vv)[..., :hd]
    want = o.reshape(B * Tq, H * hd) * torch.sigmoid(gate)
    got = out.view(B * Tq, H, hd_pad)[..., :hd].reshape(B * Tq, H * hd)
    _assert_close(got, want, tol=2e-2)
    if hd_pad > hd:
        assert out.view(B * Tq, H, hd_pad)[..., hd:].abs().max().item() == 0.0


@pytest.mark.parametrize("C_,T", [(32, 1700), (64, 900), (128, 300), (256, 130), (2048, 21)])
def test_convnext_mix(eng, C_, T):
    from smalltts_b200 import _cabi

    torch.manual_seed(8)
    B = 2
    x = torch.randn(B, T, C_, device="cuda")
    nw, fw = 1 + 0.1 * torch.randn(C_, device="cuda"), 1 + 0.1 * torch.randn(C_, device="cuda")
    cw, cb = torch.randn(C_, 7, device="cuda") * 0.4, torch.randn(C_, device="cuda")
    gamma = 0.1 + 0.1 * torch.rand(C_, device="cuda")
    y = torch.zeros_like(x)
    a = torch.zeros(B, T, C_, device="cuda", dtype=torch.bfloat16)
    rc = _cabi.lib().stts_test_convnext_mix(eng._h, _p(x), B, T, C_, _p(nw), _p(cw), _p(cb), _p(gamma), _p(fw), _p(y),
                                            _p(a))
    _cabi.check(rc, eng._h)
    torch.cuda.synchronize()
    xn = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-5) * nw
    conv = torch.nn.functional.conv1d(torch.nn.functional.pad(xn.transpose(1, 2), (6, 0)), cw[:, None, :], cb, groups=C_)
    want_y = x + gamma * conv.transpose(1, 2)
    want_a = want_y * torch.rsqrt(want_y.pow(2).mean(-1, keepdim=True) + 1e-5) * fw
    _assert_close(y, want_y, tol=1e-5)
    _assert_close(a, want_a, tol=1e-2)


@pytest.mark.parametrize("C_,T", [(32, 1000), (64, 700), (64, 128), (32, 57), (64, 19000), (32, 30011)])
def test_convnext_fused(eng, C_, T):
    """Whole ConvNeXt layer (hf:284-297) in one kernel vs torch fp32 on bf16-rounded weights."""
    from smalltts_b200 import _cabi

    torch.manual_seed(9)
    B = 3
    x = torch.randn(B, T, C_, device="cuda")
    nw, fw = 1 + 0.1 * torch.randn(C_, device="cuda"), 1 + 0.1 * torch.randn(C_, device="cuda")
    cw, cb = torch.randn(C_, 7, device="cuda") * 0.4, torch.randn(C_, device="cuda")
    gamma, fgamma = 0.1 + 0.1 * torch.rand(C_, device="cuda"), 0.1 + 0.1 * torch.rand(C_, device="cuda")
    w1 = _rand_bf16(4 * C_, C_, scale=C_ ** -0.5)
    w2 = (torch.randn(C_, 4 * C_, device="cuda") * (4 * C_) ** -0.5).to(torch.float16)  # fused kernel: fp16 W2
    b1, b2 = torch.randn(4 * C_, device="cuda") * 0.3, torch.randn(C_, device="cuda") * 0.3
    out = torch.zeros_like(x)
    out16 = torch.zeros(B, T, C_, device="cuda", dtype=torch.bfloat16)
    rc = _cabi.lib().stts_test_convnext_fused(eng._h, _p(x), B, T, C_, _p(nw), _p(cw), _p(cb), _p(gamma), _p(fw),
                                              _p(w1), _p(b1), _p((w2 * 0.5).contiguous()), _p(b2), _p(fgamma), _p(out), _p(out16))  # kernel takes 0.5*W2
    _cabi.check(rc, eng._h)
    torch.cuda.synchronize()
    xn = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-5) * nw
    conv = torch.nn.functional.conv1d(torch.nn.functional.pad(xn.transpose(1, 2), (6, 0)), cw[:, None, :], cb, groups=C_)
    y = x + gamma * conv.transpose(1, 2)
    a = (y * torch.rsqrt(y.pow(2).mean(-1, keepdim=True) + 1e-5) * fw).to(torch.bfloat16).float()
    h = torch.nn.functional.gelu(a @ w1.float().t() + b1).to(torch.float16).float()
    want = y + fgamma * (h @ w2.float().t() + b2)
    _assert_close(out, want, tol=2e-3)
    _assert_close(out16, want, tol=1e-2)


@pytest.mark.parametrize("C_", [32, 64])
def test_convnext_fused_long_tile_sequences_are_deterministic(eng, C_):
    """Every CTA of the persistent fused layer walks 13 tiles here (the pipeline's buffer rotations and barrier parities go
    through several full cycles, which the short cases above do not reach): result vs torch, and bit-identical run to
    run -- the warp roles hand tiles over through shared memory, so a protocol slip shows up as run-to-run noise."""
    from smalltts_b200 import _cabi

    torch.manual_seed(11)
    B, T = 2, 148 * 13 * 64 - 37  # 2 x 962 tiles of 128 rows on 148 CTAs, ragged last tile
    x = torch.randn(B, T, C_, device="cuda")
    nw, fw = 1 + 0.1 * torch.randn(C_, device="cuda"), 1 + 0.1 * torch.randn(C_, device="cuda")
    cw, cb = torch.randn(C_, 7, device="cuda") * 0.4, torch.randn(C_, device="cuda")
    gamma, fgamma = 0.1 + 0.1 * torch.rand(C_, device="cuda"), 0.1 + 0.1 * torch.rand(C_, device="cuda")
    w1 = _rand_bf16(4 * C_, C_, scale=C_ ** -0.5)
    w2 = (torch.randn(C_, 4 * C_, device="cuda") * (4 * C_) ** -0.5).to(torch.float16)
    b1, b2 = torch.randn(4 * C_, device="cuda") * 0.3, torch.randn(C_, device="cuda") * 0.3
    w2h = (w2 * 0.5).contiguous()
    outs = []
    for _ in range(3):
        out = torch.zeros_like(x)
        rc = _cabi.lib().stts_test_convnext_fused(eng._h, _p(x), B, T, C_, _p(nw), _p(cw), _p(cb), _p(gamma), _p(fw),
                                                  _p(w1), _p(b1), _p(w2h), _p(b2), _p(fgamma), _p(out), None)
        _cabi.check(rc, eng._h)
        torch.cuda.synchronize()
        outs.append(out)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    xn = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-5) * nw
    conv = torch.nn.functional.conv1d(torch.nn.functional.pad(xn.transpose(1, 2), (6, 0)), cw[:, None, :], cb, groups=C_)
    y = x + gamma * conv.transpose(1, 2)
    a = (y * torch.rsqrt(y.pow(2).mean(-1, keepdim=True) + 1e-5) * fw).to(torch.bfloat16).float()
    h = torch.nn.functional.gelu(a @ w1.float().t() + b1).to(torch.float16).float()
    want = y + fgamma * (h @ w2.float().t() + b2)
    _assert_close(outs[0], want, tol=2e-3)


@pytest.mark.parametrize("M", [128, 1000, 37, 40000])
def test_ffn_fused(eng, M):
    """ConvNeXt feed-forward for C = 128 in one kernel (hidden activation in TMEM / shared memory only) vs torch fp32
    on the same bf16 / fp16-rounded operands."""
    from smalltts_b200 import _cabi

    torch.manual_seed(12)
    C_ = 128
    a = _rand_bf16(M, C_)
    y = torch.randn(M, C_, device="cuda")
    w1 = _rand_bf16(4 * C_, C_, scale=C_ ** -0.5)
    w2 = (torch.randn(C_, 4 * C_, device="cuda") * (4 * C_) ** -0.5).to(torch.float16)
    b1, b2 = torch.randn(4 * C_, device="cuda") * 0.3, torch.randn(C_, device="cuda") * 0.3
    fgamma = 0.1 + 0.1 * torch.rand(C_, device="cuda")
    out = torch.zeros_like(y)
    out16 = torch.zeros(M, C_, device="cuda", dtype=torch.bfloat16)
    rc = _cabi.lib().stts_test_ffn_fused(eng._h, _p(a), _p(y), M, C_, _p(w1), _p(b1), _p((w2 * 0.5).contiguous()), _p(b2),
                                         _p(fgamma), _p(out), _p(out16))  # kernel takes 0.5*W2
    _cabi.check(rc, eng._h)
    torch.cuda.synchronize()
    h = torch.nn.functional.gelu(a.float() @ w1.float().t() + b1).to(torch.float16).float()
    want = y + fgamma * (h @ w2.float().t() + b2)
    _assert_close(out, want, tol=2e-3)
    _assert_close(out16, want, tol=1e-2)

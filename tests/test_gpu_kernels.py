"""GPU parity tests, kernel by kernel, through the C ABI (ctypes) against the CPU oracle.

Tolerances (BASELINE.json north_star): integer / index / placement work bit-exact; floating point
max|delta|/max|ref| <= 1e-2 and cosine >= 0.9999 against the fp32 oracle — single kernels are held to
much tighter bounds here (bf16 operands, fp32 accumulation: a few 1e-3).
"""

import math

import numpy as np
import pytest
import torch

from oracle import fusion_oracle as FO
from oracle import vision_oracle as VO

pytestmark = pytest.mark.gpu
IMG = 248056


@pytest.fixture(scope="module")
def L():
    from llm_quest_b200 import _lib

    _lib.lib()
    return _lib


def dev(t):
    return t.cuda()


def bf(t):
    return t.to(torch.bfloat16)


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def check_close(got, ref, tol=1e-2, cos=0.9999, what=""):
    got, ref = got.float().cpu(), ref.float().cpu()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert torch.isfinite(got).all(), f"{what}: non-finite output"
    e, c = VO.max_norm_err(got, ref), VO.cosine(got, ref)
    print(f"{what}: max_norm_err={e:.3e} cosine={c:.7f}")
    assert e <= tol and c >= cos, f"{what}: err {e:.3e} cos {c:.6f}"


# ------------------------------------------------------------------------------------------------
# GEMM + epilogues
# ------------------------------------------------------------------------------------------------
GEMM_SHAPES = [(128, 256, 64), (300, 768, 768), (1000, 2304, 768), (257, 100, 768), (8, 1024, 3072), (4096, 3072, 768)]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_bias_f32(L, M, N, K):
    a, w, b = bf(rnd(M, K, seed=1)), bf(rnd(N, K, seed=2, scale=0.05)), rnd(N, seed=3)
    out = torch.full((M, N), float("nan"), device="cuda")
    L.gemm(dev(a), dev(w), L.VF_EPI_BIAS_F32, out, bias=dev(b))
    check_close(out, a.float() @ w.float().t() + b, tol=2e-3, what=f"gemm_f32 {M}x{N}x{K}")


def test_gemm_no_bias_bf16_and_strided_a(L):
    M, N, K = 200, 512, 256
    big = bf(rnd(M, 2 * K, seed=4))
    a = big[:, :K]  # row stride 2K
    w = bf(rnd(N, K, seed=5, scale=0.05))
    out = torch.zeros((M, N), dtype=torch.bfloat16, device="cuda")
    L.gemm(dev(big)[:, :K], dev(w), L.VF_EPI_BIAS_BF16, out)
    check_close(out, a.float() @ w.float().t(), tol=6e-3, what="gemm_bf16 strided A")


def test_gemm_residual_inplace(L):
    M, N, K = 777, 768, 3072
    a, w, b = bf(rnd(M, K, seed=6)), bf(rnd(N, K, seed=7, scale=0.02)), rnd(N, seed=8)
    res = rnd(M, N, seed=9)
    x = dev(res.clone())
    L.gemm(dev(a), dev(w), L.VF_EPI_BIAS_RES_F32, x, bias=dev(b), res=x)
    check_close(x, a.float() @ w.float().t() + b + res, tol=2e-3, what="gemm residual in-place")


@pytest.mark.parametrize("mode,fn", [("tanh", VO.gelu_tanh), ("erf", VO.gelu_erf)])
def test_gemm_gelu(L, mode, fn):
    M, N, K = 640, 3072, 768
    a, w, b = bf(rnd(M, K, seed=10)), bf(rnd(N, K, seed=11, scale=0.05)), rnd(N, seed=12)
    out = torch.zeros((M, N), dtype=torch.bfloat16, device="cuda")
    L.gemm(dev(a), dev(w), L.VF_EPI_GELU_TANH_BF16 if mode == "tanh" else L.VF_EPI_GELU_ERF_BF16, out, bias=dev(b))
    check_close(out, fn(a.float() @ w.float().t() + b), tol=6e-3, what=f"gemm gelu_{mode}")


def test_gemm_qkv_rope(L):
    nh, nw, frames, B, H = 5, 3, 2, 3, 2
    n, D = nh * nw, 128
    S = frames * n
    M = B * S
    a, w, b = bf(rnd(M, D, seed=13)), bf(rnd(3 * D, D, seed=14, scale=0.08)), rnd(3 * D, seed=15)
    cos, sin = VO.axial_rope_tables(10_000, 64, nh, nw)
    out = torch.zeros((M, 3 * D), dtype=torch.bfloat16, device="cuda")
    L.gemm(dev(a), dev(w), L.VF_EPI_QKV_ROPE_BF16, out, bias=dev(b),
           rope=(dev(cos[:, :32].contiguous()), dev(sin[:, :32].contiguous()), n, 2 * D))
    qkv = (a.float() @ w.float().t() + b).view(B, S, 3, H, 64)
    q, k, v = (qkv[:, :, i].transpose(1, 2) for i in range(3))
    cs, sn = cos.repeat(frames, 1), sin.repeat(frames, 1)
    q, k = VO.rotate_half_apply(q, cs, sn), VO.rotate_half_apply(k, cs, sn)
    ref = torch.stack([t.transpose(1, 2) for t in (q, k, v)], dim=2).reshape(M, 3 * D)
    check_close(out, ref, tol=6e-3, what="gemm qkv+rope")


def test_gemm_row_remap_and_scatter(L):
    M, N, K = 30, 256, 128
    a, w = bf(rnd(M, K, seed=16)), bf(rnd(N, K, seed=17, scale=0.1))
    ref = a.float() @ w.float().t()
    fused = torch.zeros((3, 25, N), device="cuda")  # 3 samples x (10 vision + 15 text) rows
    L.gemm(dev(a), dev(w), L.VF_EPI_BIAS_F32, fused.view(-1, N), grp_rows=10, grp_stride=25, row_off=0)
    got = fused.cpu()
    check_close(got[:, :10].reshape(M, N), ref, tol=2e-3, what="gemm row remap")
    assert (got[:, 10:] == 0).all()
    dst = torch.full((M,), -1, dtype=torch.int32)
    perm = torch.randperm(64, generator=torch.Generator().manual_seed(1))[:M].to(torch.int32)
    dst[: M - 3] = perm[: M - 3]
    out = torch.zeros((64, N), dtype=torch.bfloat16, device="cuda")
    L.gemm(dev(a), dev(w), L.VF_EPI_SCATTER_BF16, out, dst_rows=dev(dst))
    got = out.float().cpu()
    exp = torch.zeros(64, N)
    exp[dst[: M - 3].long()] = ref[: M - 3]
    check_close(got, exp, tol=6e-3, what="gemm scatter")
    untouched = torch.ones(64, dtype=torch.bool)
    untouched[dst[: M - 3].long()] = False
    assert (got[untouched] == 0).all()


def _fold(w, b, gamma, beta):
    """host side of the folded LayerNorm (qwen3_5_vision_model._fold_ln restated for the test)"""
    wf = bf(w * gamma[None, :])
    return wf, (b + w @ beta).contiguous(), wf.float().sum(1).contiguous()


@pytest.mark.parametrize("M", [777, 1024])
@pytest.mark.parametrize("with_shift", [False, True])
def test_gemm_folded_layernorm_producer(L, M, with_shift):
    """bias_res_f32 with ln_out: bf16 copy of the rows minus their shift + per-32-column partial (sum, sum of squares)
    of the shifted rows; the fp32 rows themselves are not shifted."""
    N, K = 768, 256
    a, w, b = bf(rnd(M, K, seed=60)), bf(rnd(N, K, seed=61, scale=0.05)), rnd(N, seed=62)
    res = rnd(M, N, seed=63) + 0.3
    shift = (rnd(M, seed=64) * 3.0) if with_shift else torch.zeros(M)
    x = dev(res.clone())
    xb = torch.zeros((M, N), dtype=torch.bfloat16, device="cuda")
    stat = torch.full((N // 32, M, 2), float("nan"), device="cuda")
    L.gemm(dev(a), dev(w), L.VF_EPI_BIAS_RES_F32, x, bias=dev(b), res=x, ln_out=(xb, stat, dev(shift) if with_shift else None))
    ref = a.float() @ w.float().t() + b + res
    check_close(x, ref, tol=2e-3, what="producer fp32 rows")
    xs = x.cpu() - shift[:, None]
    assert torch.equal(xb.cpu(), bf(xs)), "bf16 copy is not the rounding of the shifted fp32 row"
    blocks = xs.view(M, N // 32, 32)
    torch.testing.assert_close(stat[:, :, 0].cpu().t(), blocks.sum(-1), rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(stat[:, :, 1].cpu().t(), (blocks * blocks).sum(-1), rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("N", [768, 800])
@pytest.mark.parametrize("with_ln", [False, True])
def test_gemm_residual_tma_epilogue_matches_staged_epilogue_bit_exact(L, with_ln, N):
    """Large M takes the CTA-pair kernel whose residual epilogue runs through TMA (32x32 blocks updated in place in
    shared memory); small M takes the single-CTA kernel with the staged register epilogue. Same rows, same bits —
    fp32 rows, shifted bf16 copy and LayerNorm partial sums — and a ragged last row block (4500 = 35 * 128 + 20)."""
    M, K = 4500, 320          # N = 800: the last 256-wide column tile is ragged (blocks right of N are skipped / clipped)
    a, w, b = dev(bf(rnd(M, K, seed=90))), dev(bf(rnd(N, K, seed=91, scale=0.05))), dev(rnd(N, seed=92))
    res = rnd(M, N, seed=93) + 0.25
    shift = dev(rnd(M, seed=94) * 0.5 + 0.25)
    ref = a.float().cpu() @ w.float().cpu().t() + b.cpu() + res

    def run(rows):
        x = dev(res[rows].clone())
        n = x.shape[0]
        xb = torch.zeros((n, N), dtype=torch.bfloat16, device="cuda")
        stat = torch.full((N // 32, n, 2), float("nan"), device="cuda")
        L.gemm(a[rows], w, L.VF_EPI_BIAS_RES_F32, x, bias=b, res=x, ln_out=(xb, stat, shift[rows].contiguous()) if with_ln else None)
        return x.cpu(), xb.cpu(), stat.cpu()

    big = run(slice(0, M))
    check_close(big[0], ref, tol=2e-3, what="TMA residual epilogue vs fp32 oracle")
    for lo in (0, 2048, 4096):
        hi = min(lo + 1024, M)
        small = run(slice(lo, hi))                       # 8 row blocks: 128-wide tiles, staged epilogue
        assert torch.equal(big[0][lo:hi], small[0]), f"fp32 rows differ between the two epilogues at {lo}"
        if with_ln:
            assert torch.equal(big[1][lo:hi], small[1]), "bf16 copies differ"
            assert torch.equal(big[2][:, lo:hi], small[2]), "LayerNorm partial sums differ"
    if with_ln:
        xs = big[0] - shift.cpu()[:, None]
        assert torch.equal(big[1], bf(xs))
        torch.testing.assert_close(big[2][:, :, 0].t(), xs.view(M, N // 32, 32).sum(-1), rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("mode", ["tanh", "erf", "qkv"])
def test_gemm_folded_layernorm_consumer(L, mode):
    """producer -> consumer chain equals LayerNorm followed by the plain GEMM (oracle: fp32 LN + Linear)."""
    nh, nw, B = 6, 5, 9
    n = nh * nw
    M, D = B * n, 768
    N = 3 * D if mode == "qkv" else 1024
    a, w0, b0 = bf(rnd(M, 128, seed=70)), bf(rnd(D, 128, seed=71, scale=0.1)), rnd(D, seed=72)
    res = rnd(M, D, seed=73) * 2.0 + 0.5          # non-zero row means: the mean term must cancel
    gamma, beta = 1.0 + 0.2 * rnd(D, seed=74), 0.1 * rnd(D, seed=75)
    w, b = rnd(N, D, seed=76, scale=0.05), rnd(N, seed=77)
    x = dev(res.clone())
    xb = torch.empty((M, D), dtype=torch.bfloat16, device="cuda")
    stat = torch.empty((D // 32, M, 2), device="cuda")
    L.gemm(dev(a), dev(w0), L.VF_EPI_BIAS_RES_F32, x, bias=dev(b0), res=x, ln_out=(xb, stat))
    wf, bfold, cs = _fold(w, b, gamma, beta)
    xr = x.cpu()
    rows = torch.empty((M, 2), device="cuda")
    L.ln_row_stats(stat, D, 1e-6, rows)
    torch.testing.assert_close(rows[:, 0].cpu(), xr.mean(1), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(rows[:, 1].cpu(), (xr.var(1, unbiased=False) + 1e-6).rsqrt(), rtol=1e-4, atol=1e-5)
    # variant 1 = the Part-1 LayerNorm (eps added to the std)
    rows1 = torch.empty((M, 2), device="cuda")
    L.ln_row_stats(stat, D, 1e-5, rows1, variant=1)
    torch.testing.assert_close(rows1[:, 1].cpu(), 1.0 / (xr.var(1, unbiased=False).sqrt() + 1e-5), rtol=1e-4, atol=1e-5)
    lin = torch.nn.functional.layer_norm(xr, (D,), gamma, beta, 1e-6) @ w.t() + b
    out = torch.zeros((M, N), dtype=torch.bfloat16, device="cuda")
    if mode == "qkv":
        cos, sin = VO.axial_rope_tables(10_000, 64, nh, nw)
        L.gemm(xb, dev(wf), L.VF_EPI_QKV_ROPE_BF16, out, bias=dev(bfold), ln_in=(rows, dev(cs)),
               rope=(dev(cos[:, :32].contiguous()), dev(sin[:, :32].contiguous()), n, 2 * D))
        H = D // 64
        qkv = lin.view(B, n, 3, H, 64)
        q, k, v = (qkv[:, :, i].transpose(1, 2) for i in range(3))
        q, k = VO.rotate_half_apply(q, cos, sin), VO.rotate_half_apply(k, cos, sin)
        ref = torch.stack([t.transpose(1, 2) for t in (q, k, v)], dim=2).reshape(M, N)
    else:
        epi, fn = (L.VF_EPI_GELU_TANH_BF16, VO.gelu_tanh) if mode == "tanh" else (L.VF_EPI_GELU_ERF_BF16, VO.gelu_erf)
        L.gemm(xb, dev(wf), epi, out, bias=dev(bfold), ln_in=(rows, dev(cs)))
        ref = fn(lin)
    check_close(out, ref, tol=6e-3, what=f"folded layernorm -> {mode}")


@pytest.mark.parametrize("offset", [10.0, 50.0])
def test_gemm_folded_layernorm_rows_with_large_mean(L, offset):
    """Rows whose mean is 10 / 50 standard deviations away from zero (outlier rows of real checkpoints): with the row
    shift (here: the exact mean of the residual BEFORE the producing GEMM, as in the tower, where it comes from the
    previous LayerNorm point) the folded chain stays at the accuracy of LayerNorm-then-GEMM; without it the bf16
    rounding of x swamps the row's spread, which is why the shift exists."""
    M, D, N = 2048, 768, 1024
    a, w0, b0 = bf(rnd(M, 128, seed=70)), bf(rnd(D, 128, seed=71, scale=0.1)), rnd(D, seed=72)
    res = rnd(M, D, seed=73) + offset * (1.0 + 0.1 * rnd(M, 1, seed=78))          # per-row means around `offset`, sigma 1
    gamma, beta = 1.0 + 0.2 * rnd(D, seed=74), 0.1 * rnd(D, seed=75)
    w, b = rnd(N, D, seed=76, scale=0.05), rnd(N, seed=77)
    wf, bfold, cs = _fold(w, b, gamma, beta)
    errs = {}
    for use_shift in (True, False):
        x = dev(res.clone())
        shift = dev(res.mean(1).contiguous()) if use_shift else None
        xb = torch.empty((M, D), dtype=torch.bfloat16, device="cuda")
        stat = torch.empty((D // 32, M, 2), device="cuda")
        L.gemm(dev(a), dev(w0), L.VF_EPI_BIAS_RES_F32, x, bias=dev(b0), res=x, ln_out=(xb, stat, shift))
        rows = torch.empty((M, 2), device="cuda")
        L.ln_row_stats(stat, D, 1e-6, rows, shift)
        xr = x.cpu()
        if use_shift:   # the shift has been advanced to the rows' true means
            torch.testing.assert_close(shift.cpu(), xr.mean(1), rtol=1e-5, atol=1e-4)
            torch.testing.assert_close(rows[:, 1].cpu(), (xr.var(1, unbiased=False) + 1e-6).rsqrt(), rtol=1e-3, atol=1e-5)
        out = torch.zeros((M, N), dtype=torch.bfloat16, device="cuda")
        L.gemm(xb, dev(wf), L.VF_EPI_GELU_TANH_BF16, out, bias=dev(bfold), ln_in=(rows, dev(cs)))
        ref = VO.gelu_tanh(torch.nn.functional.layer_norm(xr, (D,), gamma, beta, 1e-6) @ w.t() + b)
        errs[use_shift] = VO.max_norm_err(out.float().cpu(), ref)
    print(f"rows with mean {offset} sigma: folded LayerNorm error with shift {errs[True]:.2e}, without {errs[False]:.2e}")
    assert errs[True] <= 6e-3, errs
    assert errs[False] > 2 * errs[True], "the shift should matter for such rows"


@pytest.mark.parametrize("mode,M", [("tanh", 1000), ("qkv", 1176), ("bias", 300)])
def test_gemm_folded_layernorm_statistics_in_the_consumer_same_bits(L, mode, M):
    """Small problems: the consuming GEMM adds up the producer's partial sums itself (ln_part_in) — bit for bit what the
    vf_ln_row_stats route gives, including the advanced row shift; variant 1 (eps on the std) for the Part-1 LayerNorm."""
    nh, nw = 7, 6
    n = nh * nw
    D = 768
    N = 3 * D if mode == "qkv" else 1024
    variant = 1 if mode == "bias" else 0
    eps = 1e-5 if variant else 1e-6
    a, w0, b0 = dev(bf(rnd(M, 128, seed=70))), dev(bf(rnd(D, 128, seed=71, scale=0.1))), dev(rnd(D, seed=72))
    res = rnd(M, D, seed=73) * 2.0 + 3.0
    gamma, beta = 1.0 + 0.2 * rnd(D, seed=74), 0.1 * rnd(D, seed=75)
    w, b = rnd(N, D, seed=76, scale=0.05), rnd(N, seed=77)
    wf, bfold, cs = (dev(t) for t in _fold(w, b, gamma, beta))
    cos, sin = VO.axial_rope_tables(10_000, 64, nh, nw)
    rope = (dev(cos[:, :32].contiguous()), dev(sin[:, :32].contiguous()), n, 2 * D)
    epi = {"tanh": L.VF_EPI_GELU_TANH_BF16, "qkv": L.VF_EPI_QKV_ROPE_BF16, "bias": L.VF_EPI_BIAS_BF16}[mode]
    outs, shifts = [], []
    for in_consumer in (False, True):
        x = dev(res.clone())
        shift = dev(res.mean(1).contiguous())
        xb = torch.empty((M, D), dtype=torch.bfloat16, device="cuda")
        stat = torch.empty((D // 32, M, 2), device="cuda")
        L.gemm(a, w0, L.VF_EPI_BIAS_RES_F32, x, bias=b0, res=x, ln_out=(xb, stat, shift))
        out = torch.zeros((M, N), dtype=torch.bfloat16, device="cuda")
        if in_consumer:
            ln_in = (stat, cs, eps, variant, shift)
        else:
            rows = torch.empty((M, 2), device="cuda")
            L.ln_row_stats(stat, D, eps, rows, shift, variant=variant)
            ln_in = (rows, cs)
        L.gemm(xb, wf, epi, out, bias=bfold, ln_in=ln_in, rope=rope if mode == "qkv" else None)
        outs.append(out.cpu())
        shifts.append(shift.cpu())
    assert torch.equal(outs[0], outs[1]), "the two statistics routes differ"
    assert torch.equal(shifts[0], shifts[1]), "the advanced row shifts differ"
    xr = x.cpu()
    torch.testing.assert_close(shifts[1], xr.mean(1), rtol=1e-5, atol=1e-4)
    if mode == "bias":
        ref = VO.std_layernorm(xr, gamma, beta, eps) @ w.t() + b
        check_close(outs[1], ref, tol=6e-3, what="folded Part-1 LayerNorm -> plain bf16 epilogue")


def test_gemm_folded_layernorm_rejects_bad_arguments(L):
    M, N, K = 256, 768, 768
    a, w = dev(bf(rnd(M, K, seed=80))), dev(bf(rnd(N, K, seed=81)))
    out = torch.zeros((M, N), dtype=torch.bfloat16, device="cuda")
    rows, cs = torch.zeros((M, 2), device="cuda"), torch.zeros(N, device="cuda")
    with pytest.raises(L.VFuseError):                         # epilogue without a folded-LN form
        L.gemm(a, w, L.VF_EPI_BIAS_F32, torch.zeros((M, N), device="cuda"), ln_in=(rows, cs))
    xf = torch.zeros((M, N), device="cuda")
    with pytest.raises(L.VFuseError):                         # producer needs the residual epilogue
        L.gemm(a, w, L.VF_EPI_BIAS_F32, xf, ln_out=(out, torch.zeros((N // 32, M, 2), device="cuda")))


# ------------------------------------------------------------------------------------------------
# patch embedding (im2col-free TMA gather)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,T,H,W,D", [(2, 2, 64, 96, 128), (3, 4, 448, 448, 768), (1, 2, 224, 224, 256), (2, 2, 160, 48, 128)])
def test_patch_embed3d(L, B, T, H, W, D):
    P, tp = 16, 2
    x = bf(rnd(B, 3, T, H, W, seed=20))
    w = bf(rnd(D, 3, tp, P, P, seed=21, scale=0.03))
    b = rnd(D, seed=22)
    n = (H // P) * (W // P)
    pos = rnd(n + 5, D, seed=23)
    S = (T // tp) * n
    out = torch.full((B * S, D), float("nan"), device="cuda")
    L.patch_embed(dev(x), dev(w.reshape(D, -1).contiguous()), dev(b), dev(pos), out, P, tp, S, 0)
    ref = VO.patch_embed3d(x.float(), w.float(), b) + pos[:n].repeat(T // tp, 1)[None]
    check_close(out.view(B, S, D), ref, tol=2e-3, what=f"patch_embed3d {B}x{T}x{H}x{W}")


def test_patch_embed2d_with_cls_rows(L):
    B, H, W, D, P = 3, 224, 224, 768, 16
    x, w, b = bf(rnd(B, 3, H, W, seed=24)), bf(rnd(D, 3, P, P, seed=25, scale=0.03)), rnd(D, seed=26)
    n = (H // P) * (W // P)
    pos, cls = rnd(n + 1, D, seed=27), rnd(D, seed=28)
    out = torch.full((B * (n + 1), D), float("nan"), device="cuda")
    L.patch_embed(dev(x).unsqueeze(2), dev(w.reshape(D, -1).contiguous()), dev(b), dev(pos)[1:], out, P, 1, n + 1, 1)
    L.vit_cls_pos(dev(cls), dev(pos)[0], out, B, n + 1, D)
    ref = VO.patch_embed3d(x.float().unsqueeze(2), w.float().unsqueeze(2), b)
    ref = torch.cat([cls.expand(B, 1, D), ref], dim=1) + pos[None]
    check_close(out.view(B, n + 1, D), ref, tol=2e-3, what="patch_embed2d+cls")


# ------------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,S,H", [(2, 128, 2), (3, 196, 2), (2, 197, 12), (1, 784, 3), (2, 24, 1), (1, 1000, 2), (1, 3136, 1)])
def test_attention(L, B, S, H):
    qkv = bf(rnd(B * S, 3 * H * 64, seed=30 + S))
    out = torch.full((B * S, H * 64), float("nan"), dtype=torch.bfloat16, device="cuda")
    L.attention(dev(qkv), out, B, S, H, 1 / 8)
    q, k, v = (qkv.float().view(B, S, 3, H, 64)[:, :, i].transpose(1, 2) for i in range(3))
    ref = (torch.softmax(q @ k.transpose(-1, -2) / 8, -1) @ v).transpose(1, 2).reshape(B * S, H * 64)
    check_close(out, ref, tol=8e-3, what=f"attention B{B} S{S} H{H}")


@pytest.mark.parametrize("B,S,H", [(1, 197, 2), (3, 200, 1), (2, 300, 2), (1, 530, 1)])
def test_attention_output_store_is_clipped_at_the_sample_end(L, B, S, H):
    """The output goes out as 32-row TMA boxes: rows past the end of a sample must be clipped by the 3-D output map, not
    written into the next sample (checked by making the FIRST rows of every sample the only rows that could be hit and
    comparing them bit for bit with a run where each sample is computed alone) nor past the end of the buffer (guard rows)."""
    qkv = bf(rnd(B * S, 3 * H * 64, seed=77 + S))
    guard = 64
    buf = torch.full((B * S + guard, H * 64), 7.0, dtype=torch.bfloat16, device="cuda")
    out = buf[:B * S]
    d = dev(qkv)
    L.attention(d, out, B, S, H, 1 / 8)
    torch.cuda.synchronize()
    assert bool((buf[B * S:] == 7.0).all()), "rows behind the last sample were written"
    for b in range(B):
        alone = torch.full((S + guard, H * 64), 7.0, dtype=torch.bfloat16, device="cuda")
        L.attention(d[b * S:(b + 1) * S].contiguous(), alone[:S], 1, S, H, 1 / 8)
        assert torch.equal(alone[:S], out[b * S:(b + 1) * S]), f"sample {b} differs from the same sample computed alone"
        assert bool((alone[S:] == 7.0).all())


@pytest.mark.parametrize("B,S,H", [(80, 196, 2), (151, 197, 1), (50, 100, 3), (75, 256, 2)])
def test_attention_pair_mode(L, B, S, H):
    """S <= 256 with more (sample, head) pairs than SMs: a work item is a PAIR of (sample, head) (chains 0,1 / 2,3, two K/V
    rings); 151 x 1 leaves the last item without a second member, S = 100 gives one query tile per member."""
    qkv = bf(rnd(B * S, 3 * H * 64, seed=90 + S))
    out = torch.full((B * S, H * 64), float("nan"), dtype=torch.bfloat16, device="cuda")
    L.attention(dev(qkv), out, B, S, H, 1 / 8)
    q, k, v = (qkv.float().view(B, S, 3, H, 64)[:, :, i].transpose(1, 2) for i in range(3))
    ref = (torch.softmax(q @ k.transpose(-1, -2) / 8, -1) @ v).transpose(1, 2).reshape(B * S, H * 64)
    check_close(out, ref, tol=8e-3, what=f"attention pair mode B{B} S{S} H{H}")


def test_attention_large_scores_trigger_lazy_rescale(L):
    """Rows whose max grows tile after tile exercise the in-TMEM O rescale."""
    B, S, H = 1, 640, 1
    qkv = rnd(S, 192, seed=40)
    ramp = torch.linspace(0, 6, S)[:, None]
    qkv[:, 64:128] = qkv[:, 64:128] * 0.2 + ramp * torch.sign(qkv[:1, :64])  # keys align more and more with q
    qkv[:, :64] = qkv[:1, :64].expand(S, 64) * 2 + 0.1 * qkv[:, :64]
    qkv = bf(qkv)
    out = torch.zeros((S, 64), dtype=torch.bfloat16, device="cuda")
    L.attention(dev(qkv), out, B, S, H, 1 / 8)
    q, k, v = (qkv.float().view(1, S, 3, 1, 64)[:, :, i].transpose(1, 2) for i in range(3))
    ref = (torch.softmax(q @ k.transpose(-1, -2) / 8, -1) @ v).transpose(1, 2).reshape(S, 64)
    check_close(out, ref, tol=1e-2, what="attention rescale")


# ------------------------------------------------------------------------------------------------
# LayerNorm (+ merge gather)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("D", [128, 768, 1024])
def test_layernorm(L, variant, D):
    rows = 1003
    x, w, b = rnd(rows, D, seed=50) * 3 + 0.5, rnd(D, seed=51), rnd(D, seed=52)
    ref = torch.nn.functional.layer_norm(x, (D,), w, b, 1e-6) if variant == 0 else VO.std_layernorm(x, w, b, 1e-5)
    out = torch.zeros((rows, D), device="cuda")
    L.layernorm(dev(x), dev(w), dev(b), out, 1e-6 if variant == 0 else 1e-5, variant=variant)
    check_close(out, ref, tol=2e-5, cos=0.999999, what=f"layernorm f32 v{variant} D{D}")
    outb = torch.zeros((rows, D), dtype=torch.bfloat16, device="cuda")
    L.layernorm(dev(x), dev(w), dev(b), outb, 1e-6 if variant == 0 else 1e-5, variant=variant)
    assert torch.equal(outb.cpu(), out.cpu().to(torch.bfloat16)), "bf16 output must be the rounded fp32 result"


def test_layernorm_strided_rows(L):
    B, S, D = 5, 7, 128
    x, w, b = rnd(B, S * D, seed=53), rnd(D, seed=54), rnd(D, seed=55)
    out = torch.zeros((B, D), dtype=torch.bfloat16, device="cuda")
    L.layernorm(dev(x)[:, :D], dev(w), dev(b), out, 1e-5, variant=1)
    check_close(out, VO.std_layernorm(x[:, :D], w, b), tol=6e-3, what="layernorm strided")


def test_layernorm_merge_gather_placement_exact(L):
    """Merge indices are integer work: placement must be bit-exact."""
    B, frames, nh, nw, D = 2, 2, 4, 6, 128
    S = frames * nh * nw
    x, w, b = rnd(B * S, D, seed=56), rnd(D, seed=57), rnd(D, seed=58)
    plain = torch.zeros((B * S, D), device="cuda")
    L.layernorm(dev(x), dev(w), dev(b), plain, 1e-6)
    merged = torch.zeros((B * S // 4, 4 * D), device="cuda")
    L.layernorm(dev(x), dev(w), dev(b), merged, 1e-6, 0, 2, nh, nw)
    gi = VO.merge_gather_index(frames, nh, nw, 2)
    exp = plain.cpu().view(B, S, D)[:, gi, :].reshape(B * S // 4, 4 * D)
    assert torch.equal(merged.cpu(), exp)


# ------------------------------------------------------------------------------------------------
# RoPE / MRoPE
# ------------------------------------------------------------------------------------------------
def test_rope_apply_golden_bit_exact(L, golden_rope):
    from llm_quest_b200.common.rope import VisionRoPE

    g = golden_rope["rope2d"]
    out = VisionRoPE.apply(dev(g["x"]), dev(g["cos"]), dev(g["sin"]))
    assert torch.equal(out.cpu(), g["expected"]), "fp32 rotate-half must match the reference bit for bit"
    outb = VisionRoPE.apply(dev(bf(g["x"])), dev(g["cos"]), dev(g["sin"]))
    check_close(outb, VO.rotate_half_apply(bf(g["x"]).float(), g["cos"], g["sin"]), tol=5e-3, what="rope bf16")


def test_rope_apply_position_ids_and_partial(L):
    from llm_quest_b200.common.rope import RoPE

    cos, sin = VO.text_rope_tables(512, 10_000, 128, 0.5)  # rot 64 of 128
    x = rnd(2, 3, 9, 128, seed=60)
    pid = torch.randint(0, 512, (2, 9), generator=torch.Generator().manual_seed(3))
    out = RoPE.apply(dev(x), dev(cos), dev(sin), dev(pid))
    assert torch.equal(out.cpu(), VO.rotate_half_apply(x, cos, sin, pid))


def test_mrope_golden(L, golden_rope):
    from llm_quest_b200.common.rope import RoPE

    g = golden_rope["mrope"]
    t = g["table"]
    cos, sin = RoPE.compute_angles(t["base"], t["head_dim"], t["ctx"], rotation_factor=t["factor"])
    oc, os_ = VO.text_rope_tables(t["ctx"], t["base"], t["head_dim"], t["factor"])
    assert torch.equal(cos, oc) and torch.equal(sin, os_)
    out = RoPE.apply_mrope(dev(g["x"]), dev(cos), dev(sin), dev(g["position_ids"]), g["sections"])
    assert torch.equal(out.cpu(), g["expected"]), "fp32 MRoPE-I must match the reference bit for bit"
    w = (1.0 + g["norm_scale"]).float()
    outn = RoPE.apply_mrope(dev(g["x"]), dev(cos), dev(sin), dev(g["position_ids"]), g["sections"], norm_weight=dev(w))
    check_close(outn, g["expected_norm_mrope"], tol=1e-5, cos=0.999999, what="rmsnorm+mrope f32")
    xb = bf(g["x"])
    outb = RoPE.apply_mrope(dev(xb), dev(cos), dev(sin), dev(g["position_ids"]), g["sections"])
    check_close(outb, VO.mrope_apply(xb.float(), cos, sin, g["position_ids"], g["sections"]), tol=5e-3, what="mrope bf16")
    # interleave helper (index work): exact
    half = 32
    c3 = cos[:, :half][g["position_ids"]]
    s3 = sin[:, :half][g["position_ids"]]
    mc, ms = RoPE.interleave_mrope_coeffs(c3, s3, g["sections"])
    axes = torch.tensor(VO.mrope_slot_axes(half, g["sections"]))
    assert torch.equal(mc, c3.permute(1, 2, 3, 0)[..., torch.arange(half), axes])


# ------------------------------------------------------------------------------------------------
# position ids / early fusion (bit-exact)
# ------------------------------------------------------------------------------------------------
def test_position_ids_goldens(L, golden_fusion):
    for i, c in enumerate(golden_fusion["position_cases"]):
        ids = dev(c["ids"])
        if c["feeds"] is None:
            feeds = torch.zeros((0, 3), dtype=torch.int64)
            mask = torch.zeros_like(ids, dtype=torch.uint8)
        else:
            feeds = c["feeds"]
            mask = None if c["mask"] is None else dev(c["mask"])
        out = L.mrope_position_ids(ids, mask, IMG, feeds, 2)
        assert out.dtype == torch.int64 and torch.equal(out.cpu(), c["expected"]), f"position ids case {i}"


def test_position_ids_cfg3_full_batch(L):
    rng = np.random.default_rng(4321)
    rows = []
    for _ in range(32):
        row = []
        for i, c in enumerate([410, 410, 410, 410, 408]):
            row += list(rng.integers(0, 1000, size=c))
            if i < 4:
                row += [IMG] * 196
        rows.append(row)
    ids = np.array(rows, dtype=np.int64)
    assert ids.shape == (32, 2832)
    feeds = [[1, 28, 28]] * 4
    exp = FO.mrope_position_ids(ids, feeds)
    out = L.mrope_position_ids(dev(torch.from_numpy(ids)), None, IMG, torch.tensor(feeds), 2)
    assert torch.equal(out.cpu(), torch.from_numpy(exp)) and int(out.max()) == 2103


def test_embed_gather_scatter_goldens(L, golden_fusion):
    for c in golden_fusion["scatter_cases"]:
        tok = c["image_token_id"]
        ids, table = dev(c["ids"]), dev(c["table"])
        n_vis = c["vision"].shape[0]
        row_map, n_ph, inv = L.fuse_scan(ids, None, tok, inv_cap=n_vis)
        assert torch.equal(row_map.cpu(), c["row_map"])
        assert int(n_ph.item()) == int((c["ids"] == tok).sum())
        exp_inv = torch.full((n_vis,), -1, dtype=torch.int32)
        sel = c["row_map"] >= 0
        exp_inv[c["row_map"][sel].long()] = torch.arange(c["row_map"].numel(), dtype=torch.int32)[sel]
        assert torch.equal(inv.cpu(), exp_inv)
        for vis in (dev(c["vision"]), dev(bf(c["vision"]))):
            out = torch.zeros(c["expected"].shape, dtype=torch.bfloat16, device="cuda")
            L.embed_gather_scatter(ids, table, vis, row_map, out)
            assert torch.equal(out.cpu().view(torch.uint16), c["expected"].view(torch.uint16)), "fusion must be bit-exact"


def test_embed_gather_scatter_large_roundtrip(L):
    """Full cfg-3 size through size-independent properties: every text row equals its table row,
    placeholder row j equals vision row j, in flat order."""
    b, seq, D, vocab = 32, 2832, 1024, 5000
    g = torch.Generator().manual_seed(7)
    ids = torch.randint(0, vocab - 1, (b, seq), generator=g)
    tok = vocab - 1
    for s in range(b):
        for i in range(4):
            st = 410 * (i + 1) + 196 * i
            ids[s, st : st + 196] = tok
    table = bf(torch.randn(vocab, D, generator=g))
    n_vis = b * 4 * 196
    vis = bf(torch.randn(n_vis, D, generator=g))
    row_map, n_ph, inv = L.fuse_scan(dev(ids), None, tok, inv_cap=n_vis)
    assert int(n_ph.item()) == n_vis
    out = torch.zeros((b, seq, D), dtype=torch.bfloat16, device="cuda")
    L.embed_gather_scatter(dev(ids), dev(table), dev(vis), row_map, out)
    got = out.cpu().view(-1, D)
    flat = ids.view(-1)
    m = flat == tok
    assert torch.equal(got[~m].view(torch.uint16), table[flat[~m]].view(torch.uint16))
    assert torch.equal(got[m].view(torch.uint16), vis.view(torch.uint16))
    assert torch.equal(row_map.cpu(), torch.from_numpy(FO.scatter_row_map(ids.numpy(), None, tok)))


def test_casts(L):
    x = rnd(100003, seed=70)
    assert torch.equal(L.to_bf16(dev(x)).cpu(), x.to(torch.bfloat16))
    assert torch.equal(L.to_f32(dev(bf(x))).cpu(), bf(x).float())


def test_preprocess_u8_bit_exact(L):
    """uint8 HWC -> normalised (B, C, T, H, W): fp32 output bit-identical to the oracle (= torchvision), bf16 output
    = that value rounded once; ragged batch of two sizes, T = 1, 2, 4."""
    g = torch.Generator().manual_seed(11)
    for (B, H, W, T, mean, std) in [(3, 48, 64, 2, [0.5, 0.5, 0.5], [0.5, 0.5, 0.5]),
                                    (1, 448, 448, 2, [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]),
                                    (2, 20, 12, 4, [0.1, 0.2, 0.3], [1.0, 0.7, 0.25]), (1, 16, 16, 1, [0.0] * 3, [1.0] * 3)]:
        u8 = torch.randint(0, 256, (B, H, W, 3), generator=g, dtype=torch.uint8)
        ref = VO.preprocess_u8(u8, mean, std, T)
        got = L.preprocess_u8(dev(u8), mean, std, T, torch.float32)
        assert got.shape == ref.shape and torch.equal(got.cpu(), ref), (B, H, W, T)
        gb = L.preprocess_u8(dev(u8), mean, std, T, torch.bfloat16)
        assert torch.equal(gb.cpu(), ref.to(torch.bfloat16))


# ------------------------------------------------------------------------------------------------
# consumer side: causal GQA attention head_dim 256, in-place strided MRoPE
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,S,Hq,Hkv,causal,gate", [(2, 100, 8, 2, True, True), (1, 128, 4, 2, True, False),
                                                     (2, 300, 8, 2, True, True), (1, 257, 2, 1, False, False),
                                                     (1, 1000, 4, 4, True, True), (3, 64, 8, 2, True, True)])
def test_attention_gqa(L, B, S, Hq, Hkv, causal, gate):
    qg = bf(rnd(B * S, Hq * 512, seed=31))              # per head: 256 query columns, then 256 gate columns
    k, v = bf(rnd(B * S, Hkv * 256, seed=32)), bf(rnd(B * S, Hkv * 256, seed=33))
    out = torch.zeros((B * S, Hq * 256), dtype=torch.bfloat16, device="cuda")
    dq = dev(qg)
    L.attention_gqa(dq, dev(k), dev(v), out, B, S, Hq, Hkv, 256 ** -0.5, causal, q_col0=0, q_head_stride=512,
                    gate2d=dq if gate else None, gate_col0=256, gate_head_stride=512)
    x = qg.float().view(B, S, Hq, 512)
    q4, g4 = x[..., :256].transpose(1, 2), x[..., 256:]
    k4 = k.float().view(B, S, Hkv, 256).transpose(1, 2).repeat_interleave(Hq // Hkv, dim=1)
    v4 = v.float().view(B, S, Hkv, 256).transpose(1, 2).repeat_interleave(Hq // Hkv, dim=1)
    att = (q4 @ k4.transpose(-1, -2)) / 16.0
    if causal:
        att = att.masked_fill(torch.triu(torch.ones(S, S, dtype=torch.bool), diagonal=1), float("-inf"))
    ref = (torch.softmax(att, dim=-1) @ v4).transpose(1, 2)
    if gate:
        ref = ref * torch.sigmoid(g4)
    check_close(out, ref.reshape(B * S, Hq * 256), tol=6e-3, what=f"gqa attention B{B} S{S} Hq{Hq}/{Hkv} causal={causal} gate={gate}")


@pytest.mark.parametrize("B,S", [(1, 100), (3, 161), (2, 300)])
def test_attention_gqa_output_store_is_clipped_at_the_sample_end(L, B, S):
    """Output tiles leave as 32-row TMA boxes and gate tiles arrive the same way: rows past a sample's end must neither be
    written (guard rows; every sample equal to the same sample computed alone) nor leak gate values of the next sample."""
    Hq, Hkv, guard = 4, 2, 64
    qg = dev(bf(rnd(B * S, Hq * 512, seed=61 + S)))
    k, v = dev(bf(rnd(B * S, Hkv * 256, seed=62))), dev(bf(rnd(B * S, Hkv * 256, seed=63)))
    buf = torch.full((B * S + guard, Hq * 256), 7.0, dtype=torch.bfloat16, device="cuda")
    L.attention_gqa(qg, k, v, buf[:B * S], B, S, Hq, Hkv, 256 ** -0.5, True, q_col0=0, q_head_stride=512, gate2d=qg, gate_col0=256,
                    gate_head_stride=512)
    torch.cuda.synchronize()
    assert bool((buf[B * S:] == 7.0).all()), "rows behind the last sample were written"
    for b in range(B):
        alone = torch.full((S + guard, Hq * 256), 7.0, dtype=torch.bfloat16, device="cuda")
        sl = slice(b * S, (b + 1) * S)
        qb = qg[sl].contiguous()
        L.attention_gqa(qb, k[sl].contiguous(), v[sl].contiguous(), alone[:S], 1, S, Hq, Hkv, 256 ** -0.5, True, q_col0=0,
                        q_head_stride=512, gate2d=qb, gate_col0=256, gate_head_stride=512)
        assert torch.equal(alone[:S], buf[sl]), f"sample {b} differs from the same sample computed alone"
        assert bool((alone[S:] == 7.0).all())


def test_attention_gqa_large_scores_redo_path(L):
    """Scores that grow along the sequence force the runaway-sum redo (fresh max + O rescale) in later key tiles."""
    B, S, Hq, Hkv = 1, 320, 2, 1
    q = rnd(B * S, Hq * 256, seed=41) * 0.2
    k = rnd(B * S, Hkv * 256, seed=42) * 0.2
    k[200:] += 40.0 * q[200:201, :256].sign()          # late keys align with the queries: exponents jump by > 2^60
    q[:, :256] = q[:, :256].abs() * 8
    q, k, v = bf(q), bf(k), bf(rnd(B * S, Hkv * 256, seed=43))
    out = torch.zeros((B * S, Hq * 256), dtype=torch.bfloat16, device="cuda")
    L.attention_gqa(dev(q), dev(k), dev(v), out, B, S, Hq, Hkv, 256 ** -0.5, True)
    q4 = q.float().view(B, S, Hq, 256).transpose(1, 2)
    k4 = k.float().view(B, S, Hkv, 256).transpose(1, 2).repeat_interleave(Hq, dim=1)
    v4 = v.float().view(B, S, Hkv, 256).transpose(1, 2).repeat_interleave(Hq, dim=1)
    att = ((q4 @ k4.transpose(-1, -2)) / 16.0).masked_fill(torch.triu(torch.ones(S, S, dtype=torch.bool), diagonal=1), float("-inf"))
    assert float(att[att > -1e30].max() - att[0, 0, 250, 0]) * 1.4427 > 64      # the test really exercises the redo
    ref = (torch.softmax(att, dim=-1) @ v4).transpose(1, 2).reshape(B * S, Hq * 256)
    check_close(out, ref, tol=8e-3, what="gqa attention redo path")


def test_mrope_strided_in_place_equals_contiguous(L, golden_rope):
    """q/k heads normalised + rotated in place inside a token-major projection buffer == the [B,H,S,hd] kernel."""
    from llm_quest_b200.common.rope import RoPE

    B, H, S, hd = 2, 4, 37, 256
    cos, sin = VO.text_rope_tables(512, 10_000_000, 256, 0.25)
    pid = torch.randint(0, 400, (3, B, S), generator=torch.Generator().manual_seed(3))
    w = (1.0 + 0.1 * rnd(256, seed=8)).float()
    buf = bf(rnd(B * S, 64 + H * 512, seed=9))           # heads at columns 64 + h*512 (gate-like padding between them)
    heads = buf[:, 64:].view(B, S, H, 512)[..., :256].permute(0, 2, 1, 3).contiguous()
    ref = RoPE.apply_mrope(dev(heads), dev(cos), dev(sin), dev(pid), (11, 11, 10), norm_weight=dev(w))
    d = dev(buf)
    L.mrope_apply_heads_(d, 64, 512, B, H, S, dev(cos), dev(sin), dev(pid), (11, 11, 10), dev(w), 1e-6, 256)
    got = d[:, 64:].view(B, S, H, 512)[..., :256].permute(0, 2, 1, 3)
    assert torch.equal(got.cpu(), ref.cpu())
    untouched = d.cpu()[:, 64:].view(B, S, H, 512)[..., 256:]
    assert torch.equal(untouched, buf[:, 64:].view(B, S, H, 512)[..., 256:]) and torch.equal(d.cpu()[:, :64], buf[:, :64])


def test_position_ids_and_scatter_randomised_vs_oracle(L):
    """Bit-exact integer work on 40 random early-fusion batches: ragged placeholder runs, several feeds of different
    (t, h, w), samples without images, starved feeds, explicit masks — kernel == numpy oracle, every element."""
    rng = np.random.default_rng(2024)
    tok = 248056
    for case in range(40):
        b = int(rng.integers(1, 6))
        n_feeds = int(rng.integers(1, 4))
        feeds = [[int(rng.integers(1, 4)), 2 * int(rng.integers(1, 5)), 2 * int(rng.integers(1, 5))] for _ in range(n_feeds)]
        need = [f[0] * (f[1] // 2) * (f[2] // 2) for f in feeds]
        seq = int(sum(need) + rng.integers(n_feeds + 2, 60))
        ids = rng.integers(0, 1000, size=(b, seq)).astype(np.int64)
        for s in range(b):
            mode = rng.integers(0, 4)          # 0: all feeds, 1: no image, 2: starved last feed, 3: all feeds, shuffled gaps
            if mode == 1:
                continue
            pos = int(rng.integers(0, 3))
            for i, n in enumerate(need):
                n_put = n if not (mode == 2 and i == n_feeds - 1) else max(0, n - int(rng.integers(1, n + 1)))
                if pos + n_put > seq:
                    break
                ids[s, pos:pos + n_put] = tok
                pos += n_put + int(rng.integers(1, 4))
        use_mask = case % 3 == 0
        mask = (ids == tok) if use_mask else None
        exp = FO.mrope_position_ids(ids, feeds, mask, tok, 2)
        got = L.mrope_position_ids(dev(torch.from_numpy(ids)), None if mask is None else dev(torch.from_numpy(mask)), tok,
                                   torch.tensor(feeds), 2)
        assert np.array_equal(got.cpu().numpy(), exp), f"case {case}: feeds={feeds} b={b} seq={seq}"
        # gather/scatter placement on the same ids
        D, V = 64, 1000
        table = bf(rnd(V, D, seed=100 + case))
        n_ph = int((ids == tok).sum())
        vision = bf(rnd(max(n_ph, 1) + 3, D, seed=200 + case))
        row_map, count, _ = L.fuse_scan(dev(torch.from_numpy(ids)), None, tok)
        assert int(count.item()) == n_ph
        assert np.array_equal(row_map.cpu().numpy(), FO.scatter_row_map(ids, None, tok).reshape(-1))
        out = torch.empty((b, seq, D), dtype=torch.bfloat16, device="cuda")
        L.embed_gather_scatter(dev(torch.from_numpy(ids)), dev(table), dev(vision), row_map, out)
        flat = torch.from_numpy(ids).view(-1)
        got_e = out.cpu().view(-1, D)
        assert torch.equal(got_e[flat != tok], table[flat[flat != tok]])
        assert torch.equal(got_e[flat == tok], vision[:n_ph])


def test_error_paths_on_device(L):
    """The C ABI reports misuse instead of computing something else: misaligned rows, unsupported head_dim, wrong
    epilogue for the fused all-gather, CPU tensors."""
    a, w = dev(bf(rnd(64, 128, seed=1))), dev(bf(rnd(96, 128, seed=2)))
    out = torch.empty((64, 96), dtype=torch.bfloat16, device="cuda")
    with pytest.raises(L.VFuseError, match="fused all-gather"):
        L.gemm(a, w, L.VF_EPI_GELU_TANH_BF16, out, peer_ptrs=[out.data_ptr()])
    a_odd = dev(bf(rnd(64, 132, seed=4)))[:, :128]                         # row pitch 132 elements: not 16-byte aligned rows
    with pytest.raises(L.VFuseError, match="multiples of 8"):
        L.gemm(a_odd, w, L.VF_EPI_BIAS_BF16, out)
    q = dev(bf(rnd(64, 2 * 128, seed=3)))
    with pytest.raises(L.VFuseError, match="head_dim"):
        L.lib()  # keep the library loaded
        L.check(L.lib().vf_attention_gqa_fwd(q.data_ptr(), 256, 0, 128, q.data_ptr(), 256, q.data_ptr(), 256, out.data_ptr(), 96,
                                             None, 0, 0, 0, 1, 64, 2, 2, 128, 0.1, 1, None), "vf_attention_gqa_fwd")
    with pytest.raises(L.VFuseError, match="CPU tensor"):
        L.gemm(a.cpu(), w, L.VF_EPI_BIAS_BF16, out)


# ------------------------------------------------------------------------------------------------
# round-2 surface kernels: stand-alone GELU / zero-centred RMSNorm, Part-2 text rows, small patches, head_dim != 64
# ------------------------------------------------------------------------------------------------
def test_gelu_and_rmsnorm_goldens(L, golden_part2):
    g = golden_part2
    y = L.gelu(dev(g["gelu"]["x"]))
    torch.testing.assert_close(y.cpu(), g["gelu"]["y"], rtol=2e-6, atol=2e-6)
    x = rnd(1000, 37, seed=5) * 4                      # odd element count: vector body + scalar tail
    torch.testing.assert_close(L.gelu(dev(x)).cpu(), VO.gelu_erf(x), rtol=2e-6, atol=2e-6)
    torch.testing.assert_close(L.gelu(dev(x), tanh_form=True).cpu(), VO.gelu_tanh(x), rtol=2e-6, atol=2e-6)
    r = g["rmsnorm"]
    w = (1.0 + r["scale"]).float().contiguous()
    torch.testing.assert_close(L.rmsnorm_zc(dev(r["x"]), dev(w), 1e-6).cpu(), r["y"], rtol=2e-6, atol=2e-6)
    got = L.rmsnorm_zc(dev(r["x"].to(torch.bfloat16)), dev(w), 1e-6)
    assert got.dtype == torch.bfloat16
    assert (got.float().cpu() - r["y_bf16"].float()).abs().max() <= 2.0 ** -7 * r["y_bf16"].float().abs().max()


def test_embed_pos_concat_golden_exact(L, golden_part2):
    """Part-2 fusion: token + position embeddings written into rows n_vis.. of the fused buffer — exact fp32 adds, the
    vision rows of the buffer untouched."""
    p2 = golden_part2["part2"]
    b, seq = p2["ids"].shape
    n_vis = p2["vision"].shape[1]
    fused = torch.zeros((b, n_vis + seq, 64), device="cuda")
    fused[:, :n_vis] = dev(p2["vision"])
    L.embed_pos_concat(dev(p2["ids"]), dev(p2["tok"]), dev(p2["pos"]), fused, n_vis)
    assert torch.equal(fused.cpu(), p2["fused"])
    with pytest.raises(L.VFuseError):                              # more tokens than position embeddings
        L.embed_pos_concat(dev(torch.zeros((1, 40), dtype=torch.long)), dev(p2["tok"]), dev(p2["pos"]), torch.zeros((1, 45, 64), device="cuda"), 5)


@pytest.mark.parametrize("B,S,H,hd", [(3, 65, 8, 32), (2, 197, 3, 32), (1, 300, 2, 128), (2, 50, 4, 96)])
def test_attention_other_head_dims(L, B, S, H, hd):
    """vf_attention_fwd_hd: head dims the tcgen05 kernel is not built for (TINY_VIT_CONFIG: 8 heads of 32, S = 65)."""
    qkv = bf(rnd(B * S, 3 * H * hd, seed=40))
    out = torch.full((B * S, H * hd), float("nan"), dtype=torch.bfloat16, device="cuda")
    L.attention(dev(qkv), out, B, S, H, hd ** -0.5, head_dim=hd)
    q, k, v = (qkv.float().view(B, S, 3, H, hd)[:, :, i].transpose(1, 2) for i in range(3))
    ref = (torch.softmax(q @ k.transpose(-1, -2) * hd ** -0.5, -1) @ v).transpose(1, 2).reshape(B * S, H * hd)
    check_close(out, ref, tol=4e-3, what=f"attention head_dim {hd} S={S}")
    with pytest.raises(L.VFuseError):
        L.attention(dev(bf(rnd(4, 3 * 20, seed=1))), torch.zeros((4, 20), dtype=torch.bfloat16, device="cuda"), 1, 4, 1, 1.0, head_dim=20)


def test_tiny_vit_dims_golden(golden_part2):
    """A ViT at the TINY_VIT_CONFIG dims (4x4 patches -> vf_im2col_patches + GEMM, head_dim 32 -> CUDA-core attention,
    10-class head -> scalar epilogue) against the live reference's output."""
    from llm_quest_b200.multimodal.vision_transformer.vit_model import ViTModel

    tv = golden_part2["tiny_vit"]
    m = ViTModel(tv["cfg"])
    m.load_state_dict({k: v.float() for k, v in tv["state_dict"].items()})
    m = m.cuda().eval()
    img = tv["images"].float().cuda()
    with torch.inference_mode():
        check_close(m(img, output_hidden_states=True), tv["hidden"], what="tiny-config ViT hidden vs reference")
        check_close(m(img), tv["logits"], what="tiny-config ViT logits vs reference")
        emb = m.patch_embedding(img)                                # PatchEmbedding2D.forward on its own
    ref_emb = VO.patch_embed3d(tv["images"].float().unsqueeze(2), tv["state_dict"]["patch_embedding.conv_proj.weight"].float().unsqueeze(2),
                               tv["state_dict"]["patch_embedding.conv_proj.bias"].float())
    ref_emb = torch.cat([tv["state_dict"]["patch_embedding.cls_token"].float().expand(4, -1, -1), ref_emb], 1)
    check_close(emb, ref_emb, tol=2e-3, what="PatchEmbedding2D with 4x4 patches")

"""Stand-alone timing + parity of vf_gemm_bf16 on the cfg-2 tower shapes (development A/B tool).

    python tools/gemm_bench.py            # all four tower GEMMs, M = 64*784
"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llm_quest_b200 import _lib as L  # noqa: E402


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def err(got, ref):
    return ((got.float() - ref).abs().max() / ref.abs().max()).item()


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 64 * 784
    D, F = 768, 3072
    g = torch.Generator(device="cuda").manual_seed(0)
    rnd = lambda *s, sc=1.0: torch.randn(*s, device="cuda", generator=g) * sc
    rows = slice(M - 300, M)   # parity on the last rows (covers a ragged last m-block)

    # --- QKV + RoPE (N=2304, K=768)
    a, w, b = rnd(M, D).bfloat16(), rnd(3 * D, D, sc=0.03).bfloat16(), rnd(3 * D)
    n = 784
    ang = torch.rand(n, 32, device="cuda", generator=g) * 6.28
    cos, sin = ang.cos().contiguous(), ang.sin().contiguous()
    out = torch.empty(M, 3 * D, device="cuda", dtype=torch.bfloat16)
    fn = lambda: L.gemm(a, w, L.VF_EPI_QKV_ROPE_BF16, out, bias=b, rope=(cos, sin, n, 2 * D))
    ms = timeit(fn)
    r = (a[rows].float() @ w.float().t() + b).view(-1, 36, 64)
    idx = (torch.arange(M, device="cuda")[rows] % n)
    c, s_ = cos[idx][:, None, :], sin[idx][:, None, :]
    x1, x2 = r[..., :32], r[..., 32:]
    rot = torch.cat([x1 * c - x2 * s_, x2 * c + x1 * s_], -1)
    ref = torch.cat([rot[:, :24], r[:, 24:]], 1).reshape(-1, 3 * D)
    print(f"qkv_rope  M={M} N=2304 K=768 : {ms*1e3:7.1f} us {2*M*2304*768/ms/1e9:7.1f} TF err={err(out[rows], ref):.2e}", flush=True)

    # --- proj (N=768, K=768) with fp32 residual in place
    w2, b2 = rnd(D, D, sc=0.03).bfloat16(), rnd(D)
    x = rnd(M, D)
    x0 = x.clone()
    fn = lambda: L.gemm(a, w2, L.VF_EPI_BIAS_RES_F32, x, bias=b2, res=x)
    x.copy_(x0); fn(); torch.cuda.synchronize()
    e_ = err(x[rows], x0[rows] + a[rows].float() @ w2.float().t() + b2)
    ms = timeit(fn)
    print(f"proj_res  M={M} N=768  K=768 : {ms*1e3:7.1f} us {2*M*768*768/ms/1e9:7.1f} TF err={e_:.2e}", flush=True)

    # --- lin1 + tanh-GELU (N=3072, K=768)
    w3, b3 = rnd(F, D, sc=0.03).bfloat16(), rnd(F)
    gout = torch.empty(M, F, device="cuda", dtype=torch.bfloat16)
    fn = lambda: L.gemm(a, w3, L.VF_EPI_GELU_TANH_BF16, gout, bias=b3)
    ms = timeit(fn)
    ref = torch.nn.functional.gelu(a[rows].float() @ w3.float().t() + b3, approximate="tanh")
    print(f"lin1_gelu M={M} N=3072 K=768 : {ms*1e3:7.1f} us {2*M*F*D/ms/1e9:7.1f} TF err={err(gout[rows], ref):.2e}", flush=True)

    # --- lin2 (N=768, K=3072) with residual
    w4, b4 = rnd(D, F, sc=0.02).bfloat16(), rnd(D)
    fn = lambda: L.gemm(gout, w4, L.VF_EPI_BIAS_RES_F32, x, bias=b4, res=x)
    x.copy_(x0); fn(); torch.cuda.synchronize()
    e_ = err(x[rows], x0[rows] + gout[rows].float() @ w4.float().t() + b4)
    ms = timeit(fn)
    print(f"lin2_res  M={M} N=768  K=3072: {ms*1e3:7.1f} us {2*M*F*D/ms/1e9:7.1f} TF err={e_:.2e}", flush=True)

    # --- folded LayerNorm: the same four GEMMs with the producer / consumer epilogue extras
    xb = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
    stat = torch.empty(D // 32, M, 2, device="cuda")
    rs = torch.empty(M, 2, device="cuda")
    cs3, csq = rnd(F), rnd(3 * D)
    for name, fn, fl in [
        ("proj_res+ln_out ", lambda: L.gemm(a, w2, L.VF_EPI_BIAS_RES_F32, x, bias=b2, res=x, ln_out=(xb, stat)), 2 * M * D * D),
        ("lin2_res+ln_out ", lambda: L.gemm(gout, w4, L.VF_EPI_BIAS_RES_F32, x, bias=b4, res=x, ln_out=(xb, stat)), 2 * M * F * D),
        ("ln_row_stats    ", lambda: L.ln_row_stats(stat, D, 1e-6, rs), 0),
        ("lin1_gelu+ln_in ", lambda: L.gemm(a, w3, L.VF_EPI_GELU_TANH_BF16, gout, bias=b3, ln_in=(rs, cs3)), 2 * M * F * D),
        ("qkv_rope+ln_in  ", lambda: L.gemm(a, w, L.VF_EPI_QKV_ROPE_BF16, out, bias=b, rope=(cos, sin, n, 2 * D), ln_in=(rs, csq)), 2 * M * 3 * D * D),
    ]:
        x.copy_(x0)
        ms = timeit(fn)
        print(f"{name} M={M}: {ms*1e3:7.1f} us {fl/ms/1e9:7.1f} TF", flush=True)

    # --- cuBLAS reference points (plain bf16 GEMMs, no epilogue) for the same shapes
    for (N, K, name) in [(2304, 768, "cublas qkv "), (768, 768, "cublas proj"), (3072, 768, "cublas lin1"), (768, 3072, "cublas lin2")]:
        aa = rnd(M, K).bfloat16(); ww = rnd(N, K, sc=0.03).bfloat16()
        ms = timeit(lambda: torch.matmul(aa, ww.t()))
        print(f"{name} N={N} K={K}: {ms*1e3:7.1f} us {2*M*N*K/ms/1e9:7.1f} TF", flush=True)


if __name__ == "__main__":
    main()

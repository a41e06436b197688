"""Where does the GEMM's MMA issuer wait? (diagnosis, vf_gemm_set_debug): per cfg-2 layer shape the share of the issuer's loop
spent waiting for operands (TMA / L2 / DRAM) and for a free accumulator (epilogue), averaged over the leader CTAs.

    python tools/gemm_issuer_probe.py
"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llm_quest_b200 import _lib as L

M, D, F = 64 * 784, 768, 3072
g = torch.Generator().manual_seed(0)
r = lambda *s: (torch.randn(*s, generator=g) * 0.05).to(torch.bfloat16).cuda()
a, gact = r(M, D), r(M, F)
wq, wo, w1, w2 = r(3 * D, D), r(D, D), r(F, D), r(D, F)
bq, bo, b1, b2 = (torch.zeros(n, device="cuda") for n in (3 * D, D, F, D))
x = torch.zeros(M, D, device="cuda")
out_q = torch.empty(M, 3 * D, dtype=torch.bfloat16, device="cuda")
out_g = torch.empty(M, F, dtype=torch.bfloat16, device="cuda")
cos = torch.ones(784, 32, device="cuda"); sin = torch.zeros(784, 32, device="cuda")
xb = torch.empty(M, D, dtype=torch.bfloat16, device="cuda"); stat = torch.empty(D // 32, M, 2, device="cuda"); shift = torch.zeros(M, device="cuda")
cases = [
    ("QKV + RoPE  (N=2304, K=768)", lambda: L.gemm(a, wq, L.VF_EPI_QKV_ROPE_BF16, out_q, bias=bq, rope=(cos, sin, 784, 2 * D)), 2.0 * M * 3 * D * D),
    ("proj + res + LN out (N=768, K=768)", lambda: L.gemm(a, wo, L.VF_EPI_BIAS_RES_F32, x, bias=bo, res=x, ln_out=(xb, stat, shift)), 2.0 * M * D * D),
    ("lin1 + GELU (N=3072, K=768)", lambda: L.gemm(a, w1, L.VF_EPI_GELU_TANH_BF16, out_g, bias=b1), 2.0 * M * F * D),
    ("lin2 + res + LN out (N=768, K=3072)", lambda: L.gemm(gact, w2, L.VF_EPI_BIAS_RES_F32, x, bias=b2, res=x, ln_out=(xb, stat, shift)), 2.0 * M * F * D),
    ("lin2 + res (N=768, K=3072)", lambda: L.gemm(gact, w2, L.VF_EPI_BIAS_RES_F32, x, bias=b2, res=x), 2.0 * M * F * D),
]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
dbg = torch.zeros(148 * 4, dtype=torch.int64, device="cuda")
for name, fn, fl in cases:
    for _ in range(3):
        fn()
    flush.zero_()                                   # cold L2, as inside the step
    torch.cuda.synchronize()
    L.lib().vf_gemm_set_debug(dbg.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record()
    torch.cuda.synchronize()
    L.lib().vf_gemm_set_debug(None)
    d = dbg.cpu().view(148, 4)[0::2].double()       # leader CTAs of the pairs issue
    d = d[d[:, 3] > 0]
    loop, ops, acc, tiles = d.mean(0).tolist()
    us = e0.elapsed_time(e1) * 1e3
    print(f"{name:38s} {us:7.1f} us {fl / us / 1e6:7.1f} TF | issuer loop {loop:9.0f} cycles, {tiles:4.1f} tiles/pair: waits for operands {100 * ops / loop:5.1f} %, "
          f"for a free accumulator {100 * acc / loop:5.1f} %, issuing / other {100 * (loop - ops - acc) / loop:5.1f} %")

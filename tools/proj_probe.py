"""Cold-L2 timing of the proj-shaped GEMM (N=K=768, M=64*784) with different epilogues (development probe)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llm_quest_b200 import _lib as L
M, D = 64 * 784, 768
g = torch.Generator(device="cuda").manual_seed(0)
rnd = lambda *s, sc=1.0: torch.randn(*s, device="cuda", generator=g) * sc
a, w, b = rnd(M, D).bfloat16(), rnd(D, D, sc=0.03).bfloat16(), rnd(D)
x = rnd(M, D)
ob = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
big = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
sink = torch.empty(1, device="cuda")
def t(fn, n=12):
    ts = []
    for _ in range(3): fn()
    for _ in range(n):
        sink.copy_(big.sum().reshape(1))     # evict L2 with CLEAN lines (a read sweep), so no write-back rides on the kernel
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts)//2] * 1e3
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which == "all":
    print("bias_bf16      (77 MB out)            : %.1f us" % t(lambda: L.gemm(a, w, L.VF_EPI_BIAS_BF16, ob, bias=b)))
    print("bias_f32       (154 MB out)           : %.1f us" % t(lambda: L.gemm(a, w, L.VF_EPI_BIAS_F32, x, bias=b)))
    print("cublas bf16 out                        : %.1f us" % t(lambda: torch.matmul(a, w.t(), out=ob)))
    xx = torch.empty_like(x)
    print("torch copy fp32 (154 in, 154 out)      : %.1f us" % t(lambda: xx.copy_(x)))
print("res in place   (154 in + 154 out, TMA ring epilogue) : %.1f us" % (
      t(lambda: L.gemm(a, w, L.VF_EPI_BIAS_RES_F32, x, bias=b, res=x))))
x2 = torch.empty_like(x)
print("res out of place                                     : %.1f us" % t(lambda: L.gemm(a, w, L.VF_EPI_BIAS_RES_F32, x2, bias=b, res=x)))
print("res = ONE row broadcast (ldr=0, always L2/L1 hits)   : %.1f us" % t(lambda: L.gemm(a, w, L.VF_EPI_BIAS_RES_F32, x2, bias=b, res=x[0:1].expand(M, D))))

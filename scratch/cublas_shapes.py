"""What cuBLAS (torch.mm / F.linear, bf16) reaches on the four GEMM shapes of a cfg-2 layer, next to libvfuse's plain-bias GEMM on the same operands."""
import sys, torch
sys.path.insert(0, "/root/repo")
from llm_quest_b200 import _lib
torch.manual_seed(0)
dev = torch.device("cuda")
M = 50176
shapes = {"qkv": (2304, 768), "proj": (768, 768), "lin1": (3072, 768), "lin2": (768, 3072), "sq8192": (8192, 8192)}
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for name, (N, K) in shapes.items():
    m = 8192 if name == "sq8192" else M
    a = torch.randn(m, K, device=dev, dtype=torch.bfloat16)
    w = torch.randn(N, K, device=dev, dtype=torch.bfloat16) * 0.03
    b = torch.randn(N, device=dev, dtype=torch.bfloat16)
    out = torch.empty(m, N, device=dev, dtype=torch.bfloat16)
    fl = 2.0 * m * N * K
    t_mm = timeit(lambda: torch.mm(a, w.t(), out=out))
    t_lin = timeit(lambda: torch.nn.functional.linear(a, w, b))
    bf = b.float()
    t_vf = timeit(lambda: _lib.gemm(a, w, _lib.VF_EPI_BIAS_BF16, out, bias=bf))
    print(f"{name:7s} M={m} N={N} K={K}: cuBLAS mm {t_mm:7.1f} us {fl/t_mm/1e6:7.1f} TF | linear+bias {t_lin:7.1f} us {fl/t_lin/1e6:7.1f} TF | vfuse bias_bf16 {t_vf:7.1f} us {fl/t_vf/1e6:7.1f} TF")

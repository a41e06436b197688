"""What torch SDPA (cuDNN / flash backends, bf16, head-major contiguous q,k,v — without the transposes the reference pays around it)
reaches on the attention shapes of the configurations, next to vf_attention_fwd on the token-major QKV buffer."""
import sys, torch
sys.path.insert(0, "/root/repo")
from llm_quest_b200 import _lib
from torch.nn.attention import SDPBackend, sdpa_kernel
torch.manual_seed(0)
dev = torch.device("cuda")
H = 12
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for name, B, S in (("cfg5:224/cfg1", 256, 196), ("cfg2", 64, 784), ("cfg5:672", 28, 1764), ("cfg3 4-frame", 32, 3136), ("cfg4", 16, 6272), ("cfg5:1344", 8, 7056)):
    qkv = torch.randn(B * S, 3 * H * 64, device=dev, dtype=torch.bfloat16)
    out = torch.empty(B * S, H * 64, device=dev, dtype=torch.bfloat16)
    q, k, v = (qkv.view(B, S, 3, H, 64)[:, :, i].transpose(1, 2).contiguous() for i in range(3))
    fl = 4.0 * B * H * S * S * 64
    res = []
    for label, be in (("cuDNN", SDPBackend.CUDNN_ATTENTION), ("flash", SDPBackend.FLASH_ATTENTION)):
        try:
            with sdpa_kernel(be):
                t = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v))
            res.append(f"{label} {t:8.1f} us {fl/t/1e6:6.1f} TF")
        except Exception as e:
            res.append(f"{label} unavailable ({type(e).__name__})")
    t = timeit(lambda: _lib.attention(qkv, out, B, S, H, 0.125))
    res.append(f"vfuse {t:8.1f} us {fl/t/1e6:6.1f} TF")
    print(f"{name:14s} B={B:3d} S={S:4d}: " + " | ".join(res))

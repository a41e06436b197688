"""Library reference point for the vision attention kernel: torch SDPA (cuDNN / flash / mem-efficient backends) and flash_attn
on the cfg-2 shape (B=64, H=12, S=784, head_dim 64, bf16, bidirectional) next to vf_attention_fwd. Development tool."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llm_quest_b200 import _lib as L
import torch.nn.functional as F
from torch.nn.attention import SDPBackend, sdpa_kernel

B, H, S, D = 64, 12, int(sys.argv[1]) if len(sys.argv) > 1 else 784, 64
qkv = torch.randn(B * S, 3 * H * D, device="cuda").to(torch.bfloat16)
out = torch.empty(B * S, H * D, device="cuda", dtype=torch.bfloat16)
fl = 4.0 * B * H * S * S * D

def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

ms = t(lambda: L.attention(qkv, out, B, S, H, D ** -0.5))
print(f"vf_attention_fwd (token-major qkv in place): {ms*1e3:8.1f} us {fl/ms/1e9:7.1f} TF")
q, k, v = (qkv.view(B, S, 3, H, D)[:, :, i].transpose(1, 2) for i in range(3))   # [B,H,S,D] strided views
ref = None
for name, be in (("cudnn", SDPBackend.CUDNN_ATTENTION), ("flash", SDPBackend.FLASH_ATTENTION), ("mem_efficient", SDPBackend.EFFICIENT_ATTENTION)):
    try:
        with sdpa_kernel(be):
            ms = t(lambda: F.scaled_dot_product_attention(q, k, v))
            o = F.scaled_dot_product_attention(q, k, v)
        err = (o.transpose(1, 2).reshape(B * S, H * D).float() - out.float()).abs().max().item()
        print(f"torch SDPA {name:14s}: {ms*1e3:8.1f} us {fl/ms/1e9:7.1f} TF  max|diff| vs ours {err:.3e}")
    except Exception as e:
        print(f"torch SDPA {name}: unavailable ({str(e)[:80]})")
try:
    from flash_attn import flash_attn_qkvpacked_func
    x = qkv.view(B, S, 3, H, D)
    ms = t(lambda: flash_attn_qkvpacked_func(x))
    print(f"flash_attn 2 qkvpacked   : {ms*1e3:8.1f} us {fl/ms/1e9:7.1f} TF")
except Exception as e:
    print("flash_attn: unavailable", str(e)[:80])

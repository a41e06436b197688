"""Stand-alone timing + parity of vf_attention_fwd (development A/B tool; run once per VF_ATTN_FLAGS value).

    VF_ATTN_FLAGS=3 python tools/attn_bench.py [B S H]...
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llm_quest_b200 import _lib as L  # noqa: E402


def run(B, S, H, iters=20):
    torch.manual_seed(0)
    qkv = torch.randn(B * S, 3 * H * 64, device="cuda").to(torch.bfloat16)
    out = torch.empty(B * S, H * 64, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        L.attention(qkv, out, B, S, H, 0.125)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        L.attention(qkv, out, B, S, H, 0.125)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    # parity on the first and last sample against torch SDPA (fp32 math)
    errs = []
    for b in (0, B - 1):
        x = qkv[b * S:(b + 1) * S].float().view(S, 3, H, 64).permute(1, 2, 0, 3)
        ref = torch.nn.functional.scaled_dot_product_attention(x[0][None], x[1][None], x[2][None])[0]
        ref = ref.permute(1, 0, 2).reshape(S, H * 64)
        got = out[b * S:(b + 1) * S].float()
        errs.append(((got - ref).abs().max() / ref.abs().max()).item())
    tf = 4.0 * B * H * S * S * 64 / (ms * 1e-3) / 1e12
    print(f"flags={os.environ.get('VF_ATTN_FLAGS', 'default')} B={B} S={S} H={H}: {ms * 1e3:8.1f} us  {tf:7.1f} TF  "
          f"max_norm_err={max(errs):.2e}", flush=True)


if __name__ == "__main__":
    args = [int(a) for a in sys.argv[1:]]
    shapes = [tuple(args[i:i + 3]) for i in range(0, len(args), 3)] or [(64, 784, 12), (16, 6272, 12), (8, 197, 12), (64, 196, 12)]
    for shp in shapes:
        run(*shp)

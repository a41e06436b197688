"""Parity + timing of vf_attention_gqa_fwd against torch SDPA (development tool).   python tools/gqa_bench.py [B S]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llm_quest_b200 import _lib as L  # noqa: E402


def run(B, S, Hq=8, Hkv=2, causal=True, gate=True, iters=10):
    g = torch.Generator(device="cuda").manual_seed(0)
    qg = (torch.randn(B * S, Hq * 512, device="cuda", generator=g)).to(torch.bfloat16)   # [q | gate] per head
    k = torch.randn(B * S, Hkv * 256, device="cuda", generator=g).to(torch.bfloat16)
    v = torch.randn(B * S, Hkv * 256, device="cuda", generator=g).to(torch.bfloat16)
    out = torch.zeros(B * S, Hq * 256, device="cuda", dtype=torch.bfloat16)
    fn = lambda: L.attention_gqa(qg, k, v, out, B, S, Hq, Hkv, 256 ** -0.5, causal, q_col0=0, q_head_stride=512,
                                 gate2d=qg if gate else None, gate_col0=256, gate_head_stride=512)
    fn()
    torch.cuda.synchronize()
    x = qg.view(B, S, Hq, 512).float()
    q4, g4 = x[..., :256].transpose(1, 2), x[..., 256:]
    k4 = k.view(B, S, Hkv, 256).float().transpose(1, 2)
    v4 = v.view(B, S, Hkv, 256).float().transpose(1, 2)
    nb = min(B, 2)
    ref = torch.nn.functional.scaled_dot_product_attention(q4[:nb], k4[:nb], v4[:nb], is_causal=causal, enable_gqa=True)
    ref = ref.transpose(1, 2)
    if gate:
        ref = ref * torch.sigmoid(g4[:nb])
    ref = ref.reshape(nb * S, Hq * 256)
    got = out[: nb * S].float()
    err = ((got - ref).abs().max() / ref.abs().max()).item()
    cos = torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0).item()
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 4.0 * B * Hq * S * S * 256 * (0.5 if causal else 1.0)
    print(f"B={B} S={S} causal={causal} gate={gate}: {ms * 1e3:8.1f} us {fl / ms / 1e9:7.1f} TF "
          f"max_norm_err={err:.2e} cos={cos:.6f}", flush=True)


if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:]]
    shapes = [tuple(a[i:i + 2]) for i in range(0, len(a), 2)] or [(2, 100), (1, 128), (2, 300), (32, 2832)]
    for B, S in shapes:
        run(B, S)
    run(2, 200, causal=False, gate=False)
    run(32, 2832, gate=False)
    run(32, 2832, causal=False, gate=False)
    run(8, 8192, gate=False)

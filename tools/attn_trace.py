"""Dump the in-kernel clock64 trace of vf_attention_fwd (trace build) as text: per key step the phase
durations of every softmax chain of block 0 and the MMA issue times.

    VF_ATTN_FLAGS=2 python tools/attn_trace.py B S H first_step n_steps
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llm_quest_b200 import _lib as L  # noqa: E402

B, S, H, first, n = (int(a) for a in sys.argv[1:6])
qkv = torch.randn(B * S, 3 * H * 64, device="cuda").to(torch.bfloat16)
out = torch.empty(B * S, H * 64, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    L.attention(qkv, out, B, S, H, 0.125)
torch.cuda.synchronize()
buf = torch.zeros(20 * n * 8, dtype=torch.int64, device="cuda")
L.lib().vf_attention_set_trace(buf.data_ptr(), first, n)
L.attention(qkv, out, B, S, H, 0.125)
torch.cuda.synchronize()
L.lib().vf_attention_set_trace(None, 0, 0)
t = buf.cpu().view(20, n, 8)
t0 = int(t[t > 0].min())
print(f"# trace B={B} S={S} H={H} steps [{first},{first + n}) flags={os.environ.get('VF_ATTN_FLAGS')} stagger={os.environ.get('VF_ATTN_STAGGER')}")
print("# softmax chain rows: tile.quarter: per step  start(rel)  wait_S  ldtm  max  exp+st_issue  st_wait+arrive | step period")
for w in range(16):
    tile, q = w // 4, w % 4
    if q not in (0,):
        continue
    prev = None
    for i in range(n):
        r = [int(v) for v in t[w, i]]
        if r[0] == 0:
            continue
        per = (r[0] - prev) if prev else 0
        prev = r[0]
        print(f"t{tile}.q{q} step {first + i:4d}: start {r[0] - t0:8d}  waitS {r[1] - r[0]:5d}  ldtm {r[2] - r[1]:4d}  max {r[3] - r[2]:4d}  "
              f"exp {r[4] - r[3]:5d}  st+arr {r[5] - r[4]:4d} | period {per:5d}")
print("# MMA issuers: tile: step: S issue at(rel)/dur, PV issue at(rel)")
for tile in range(4):
    for i in range(min(n, 6)):
        r = [int(v) for v in t[16 + tile, i]]
        if r[0] or r[2]:
            print(f"mma t{tile} step {first + i:4d}: S at {r[0] - t0 if r[0] else -1:8d} dur {r[1] - r[0]:4d}   PV at {r[2] - t0 if r[2] else -1:8d}"
                  f"   wait_sfree begin {r[4] - t0 if r[4] else -1:8d}  end {r[5] - t0 if r[5] else -1:8d}")
print("# absolute stamps of tile chains (quarter 0): step: wait_begin S_ready ld_done exp_begin exp_end arrived")
for tile in range(4):
    for i in range(min(n, 4)):
        r = [int(v) for v in t[tile * 4, i]]
        if r[0]:
            print(f"abs t{tile} step {first + i:4d}: " + " ".join(f"{v - t0:7d}" for v in r[:6]))

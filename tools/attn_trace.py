"""Dump the in-kernel clock64 trace of vf_attention_fwd (trace build) as text: per key step the phase durations of the
first softmax warp of both chains of block 0 and of the two MMA issuers.

    python tools/attn_trace.py B S H first_step n_steps
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
buf = torch.zeros(4 * n * 8, dtype=torch.int64, device="cuda")
L.lib().vf_attention_set_trace(buf.data_ptr(), first, n)
L.attention(qkv, out, B, S, H, 0.125)
torch.cuda.synchronize()
L.lib().vf_attention_set_trace(None, 0, 0)
t = buf.cpu().view(4, n, 8)
t0 = int(t[t > 0].min())
print(f"# trace B={B} S={S} H={H} steps [{first},{first + n})")
print("# softmax warp 0 of chain c: start(rel) | wait_S  ld+s_free  max  wait_token  exp(2 chunks)+wait_pv  exp+st(rest)  st_wait | period")
for c in range(2):
    prev = None
    for i in range(n):
        r = [int(v) for v in t[c, i]]
        if r[0] == 0:
            continue
        per = (r[0] - prev) if prev else 0
        prev = r[0]
        print(f"c{c} step {first + i:4d}: start {r[0] - t0:8d} | waitS {r[1] - r[0]:5d}  ld {r[2] - r[1]:4d}  max {r[3] - r[2]:4d}  token {r[4] - r[3]:5d}  "
              f"exp01+pv {r[5] - r[4]:5d}  exp23+st {r[6] - r[5]:5d}  stwait {r[7] - r[6]:4d} | period {per:5d}")
print("# issuer of chain c: at(rel) | wait_sfree  issue_S   (idle)  wait_p  issue_PV")
for c in range(2):
    for i in range(n):
        r = [int(v) for v in t[2 + c, i]]
        if r[3] == 0:
            continue
        print(f"mma{c} step {first + i:4d}: at {r[0] - t0 if r[0] else -1:8d} | wait_sfree {r[1] - r[0]:5d}  issueS {r[2] - r[1]:4d}  | at {r[3] - t0:8d} wait_p {r[4] - r[3]:5d}  issuePV {r[5] - r[4]:4d}")

"""Dump the in-kernel clock64 trace of vf_attention_fwd (trace build, VF_ATTN_FLAGS=2) as text: per key step the phase durations of
the softmax warps of block 0 (rows 0..15 = warps, chain = warp // 4) and of the MMA walkers' PV+S issue (rows 16..19 = chains).

    VF_ATTN_FLAGS=2 python tools/attn_trace.py B S H first_step n_steps [warp ...]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llm_quest_b200 import _lib as L  # noqa: E402

B, S, H, first, n = (int(a) for a in sys.argv[1:6])
warps = [int(a) for a in sys.argv[6:]] or [0, 4, 8, 12]
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
print(f"# trace B={B} S={S} H={H} steps [{first},{first + n}) of block 0; cycles")
print("# softmax warp w: start(rel) | wait_S  tmem_ld  max  exp+st_issue  st_wait+arrive | period")
tot = {}
for w in warps:
    prev = None
    for i in range(n):
        r = [int(v) for v in t[w, i]]
        if r[0] == 0:
            continue
        per = (r[0] - prev) if prev else 0
        prev = r[0]
        d = (r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3], r[5] - r[4], per)
        if per:
            tot.setdefault(w, []).append(d)
        epi = f"  | item end: o_full +{r[6] - r[5]:5d}  O store +{r[7] - r[6]:5d}" if r[6] else ""
        print(f"w{w:02d} step {first + i:4d}: start {r[0] - t0:8d} | waitS {d[0]:5d}  ld {d[1]:4d}  max {d[2]:4d}  exp {d[3]:5d}  arrive {d[4]:4d} | period {per:5d}{epi}")
for w, ds in tot.items():
    m = [sum(x[k] for x in ds) / len(ds) for k in range(6)]
    print(f"# w{w:02d} mean: waitS {m[0]:.0f} ld {m[1]:.0f} max {m[2]:.0f} exp {m[3]:.0f} arrive {m[4]:.0f} period {m[5]:.0f}")
print("# walker of chain c: at(rel) issue PV(j)+S(j+1) duration")
for c in range(4):
    for i in range(n):
        r = [int(v) for v in t[16 + c, i]]
        if r[0] == 0:
            continue
        s0 = f"  | S(0) of the item issued at {r[2] - t0:8d} (+{r[3] - r[2]})" if r[2] else ""
        print(f"mma c{c} step {first + i:4d}: at {r[0] - t0:8d} issue {r[1] - r[0]:4d}{s0}")

#!/usr/bin/env bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/attn; mkdir -p $O
cat > /tmp/one.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from llm_quest_b200 import _lib as L
B,S,H = 16, 6272, 12
qkv = torch.randn(B*S, 3*H*64, device="cuda").to(torch.bfloat16)
out = torch.empty(B*S, H*64, device="cuda", dtype=torch.bfloat16)
for _ in range(3): L.attention(qkv, out, B, S, H, 0.125)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention2 -s 2 -c 1 -o $O/a3 -f python /tmp/one.py > $O/ncu.log 2>&1; tail -3 $O/ncu.log
ls -la $O

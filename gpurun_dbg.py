import sys, torch, os
sys.path.insert(0, '.')
from llm_quest_b200 import _lib as L
B,S,H = 64,784,12
qkv = torch.randn(B*S, 3*H*64, device='cuda').to(torch.bfloat16)
out = torch.empty(B*S, H*64, device='cuda', dtype=torch.bfloat16)
for _ in range(2): L.attention(qkv, out, B, S, H, 0.125)
torch.cuda.synchronize()
os.environ['VF_ATTN_DBG']='1'
L.attention(qkv, out, B, S, H, 0.125)
torch.cuda.synchronize()

import sys, torch
sys.path.insert(0, '.')
from llm_quest_b200 import _lib as L
from oracle import vision_oracle as VO
variant = sys.argv[1]
B,T,H,W,D = 2,2,64,96,128
P,tp=16,2
g = torch.Generator().manual_seed(0)
x = torch.randn(B,3,T,H,W,generator=g).to(torch.bfloat16)
w = (torch.randn(D,3,tp,P,P,generator=g)*0.03).to(torch.bfloat16)
b = torch.randn(D,generator=g)
n=(H//P)*(W//P); S=(T//tp)*n
pos = torch.randn(n+5,D,generator=g)
out = torch.full((B*S,D), float('nan'), device='cuda')
xd, wd, bd, pd = x.cuda(), w.reshape(D,-1).contiguous().cuda(), b.cuda(), pos.cuda()
if variant == 'nopos': pd = None
if variant == 'nobias': bd = None; pd = None
L.patch_embed(xd, wd, bd, pd, out, P, tp, S, 0)
torch.cuda.synchronize()
ref = VO.patch_embed3d(x.float(), w.float(), b if bd is not None else torch.zeros(D))
if pd is not None: ref = ref + pos[:n].repeat(T//tp,1)[None]
print(variant, 'err', VO.max_norm_err(out.cpu().view(B,S,D), ref))

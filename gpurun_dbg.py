import sys, torch, os, subprocess
sys.path.insert(0, '.')
if len(sys.argv) > 1:
    from llm_quest_b200 import _lib as L
    B,S,H = 64,784,12
    qkv = torch.randn(B*S, 3*H*64, device='cuda').to(torch.bfloat16)
    out = torch.empty(B*S, H*64, device='cuda', dtype=torch.bfloat16)
    for _ in range(3): L.attention(qkv, out, B, S, H, 0.125)
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): L.attention(qkv, out, B, S, H, 0.125)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/10
    print('skew', os.environ.get('VF_ATTN_SKEW'), 'ms', round(ms,4), 'TF', round(4*B*H*S*S*64/ms/1e9,1))
else:
    for sk in [0, 250, 500, 800, 1200]:
        subprocess.run([sys.executable, __file__, 'x'], env={**os.environ, 'VF_ATTN_SKEW': str(sk)})

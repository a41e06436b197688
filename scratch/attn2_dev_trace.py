"""Trace the parked second attention design (scratch/libattn2_dev.so) — phase durations inside one softmax step."""
import ctypes as C, sys, torch
lib = C.CDLL("scratch/libattn2_dev.so")
lib.vf_attention2_launch.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_void_p]
lib.vf_attention2_set_trace.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
B, S, H, first, n = 16, 6272, 12, 60, 6
qkv = torch.randn(B * S, 3 * H * 64, device="cuda").to(torch.bfloat16)
out = torch.empty(B * S, H * 64, device="cuda", dtype=torch.bfloat16)
st = torch.cuda.current_stream().cuda_stream
run = lambda: lib.vf_attention2_launch(qkv.data_ptr(), out.data_ptr(), B, S, H, 0.125, st)
for _ in range(2): assert run() == 0
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
print(f"untraced: {e0.elapsed_time(e1)*1e3:.1f} us")
buf = torch.zeros(4 * n * 8, dtype=torch.int64, device="cuda")
lib.vf_attention2_set_trace(buf.data_ptr(), first, n)
run(); torch.cuda.synchronize()
lib.vf_attention2_set_trace(None, 0, 0)
t = buf.cpu().view(4, n, 8)
t0 = int(t[t > 0].min())
print("# chain c: start | waitS  ld  max+xchg  token | PURE exps chunk0 (32 elems) | pv wait + STTM0 + token pass | PURE exps chunk1 | period")
for c in range(2):
    prev = None
    for i in range(n):
        r = [int(v) for v in t[c, i]]
        if r[0] == 0: continue
        per = (r[0] - prev) if prev else 0; prev = r[0]
        print(f"c{c} step {first+i}: start {r[0]-t0:7d} | waitS {r[1]-r[0]:4d} ld {r[2]-r[1]:4d} max {r[3]-r[2]:4d} token {r[4]-r[3]:5d} | exp0 {r[5]-r[4]:5d} | pv+st0+pass {r[6]-r[5]:4d} | exp1 {r[7]-r[6]:5d} | period {per}")

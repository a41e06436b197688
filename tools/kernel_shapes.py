"""Per-launch CUDA-event times inside the real cfg-2 step, grouped by (kernel family, FLOPs) so that the two residual GEMMs
(proj, lin2) show up separately. Development A/B tool: run it under different VF_* switches.

    python tools/kernel_shapes.py [warm steps] [timed steps]
"""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from llm_quest_b200 import _lib  # noqa: E402
from llm_quest_b200.qwen.qwen3_5.qwen3_5_vision_model import Qwen3_5VisionModel  # noqa: E402

warm = int(sys.argv[1]) if len(sys.argv) > 1 else 15
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
torch.manual_seed(0)
m = Qwen3_5VisionModel(bench.qwen_cfg(448)).eval().cuda()
x = torch.randn(64, 3, 2, 448, 448, device="cuda").to(torch.bfloat16)
with torch.inference_mode():
    for _ in range(warm):
        m(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        m(x)
    e1.record()
    torch.cuda.synchronize()
    print("step %.3f ms (no per-launch events)" % (e0.elapsed_time(e1) / steps))
    with _lib.KernelTimer() as kt:
        for _ in range(steps):
            m(x)
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for fam, meta, a, b in kt.records:
    key = (fam, round(meta.get("flops", 0.0) / 1e9), round(meta.get("bytes", 0.0) / 1e6))
    d = agg.setdefault(key, [0, 0.0])
    d[0] += 1
    d[1] += a.elapsed_time(b)
tot = 0.0
for (fam, gf, mb), (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    tot += ms / steps
    print(f"{fam:24s} {gf:6d} GF {mb:5d} MB  x{n // steps:3d}/step  {ms / n * 1e3:8.1f} us each  {ms / steps:7.3f} ms/step")
print("sum of kernels %.3f ms/step" % tot)

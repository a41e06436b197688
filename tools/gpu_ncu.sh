#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + full captures of the dominant kernels
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep
cat > /tmp/one_step.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
import bench
from llm_quest_b200.qwen.qwen3_5.qwen3_5_vision_model import Qwen3_5VisionModel
torch.manual_seed(123)
m = Qwen3_5VisionModel(bench.qwen_cfg(448)).eval().cuda()
x = torch.randn(64, 3, 2, 448, 448).to(torch.bfloat16).cuda()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
with torch.inference_mode():
    for _ in range(n):
        m(x)
torch.cuda.synchronize()
PY
# (1) launch list of the SAME command the bench line comes from (shares must agree with bench.py's CUDA-event shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_list.log 2>&1
echo "list exit=$?"
# (2) full captures, second forward pass (layer 5): qkv, proj, lin1, lin2 GEMMs; attention; layernorm
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 72 -c 4 -o gpurun_out/prof_gemm -f python /tmp/one_step.py 2 > gpurun_out/ncu_gemm.log 2>&1
echo "gemm exit=$?"
ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 17 -c 1 -o gpurun_out/prof_attn -f python /tmp/one_step.py 2 > gpurun_out/ncu_attn.log 2>&1
echo "attn exit=$?"
ncu --set full --clock-control none -k regex:layernorm_kernel -s 1 -c 1 -o gpurun_out/prof_ln -f python /tmp/one_step.py 2 > gpurun_out/ncu_ln.log 2>&1
echo "ln exit=$?"
ncu --set full --clock-control none -k regex:ln_row_stats_kernel -s 30 -c 1 -o gpurun_out/prof_rowstats -f python /tmp/one_step.py 2 > gpurun_out/ncu_rowstats.log 2>&1
echo "rowstats exit=$?"
ls -la gpurun_out/*.ncu-rep

"""A/B on one box: LayerNorm statistics finished by vf_ln_row_stats (large-M default) vs summed by the consuming GEMM's epilogue warps."""
import sys
sys.path.insert(0, "/root/repo")
import llm_quest_b200.qwen.qwen3_5.qwen3_5_vision_model as vm
import bench
for rows in (4096, 10 ** 9, 4096, 10 ** 9):
    vm.LN_STATS_IN_CONSUMER_MAX_ROWS = rows
    print(f"== LN_STATS_IN_CONSUMER_MAX_ROWS = {rows}", flush=True)
    sys.argv = ["bench.py", "--steps", "20", "--warmup", "5", "--no-cpu", "--no-eager", "--no-u8"]
    bench.main()

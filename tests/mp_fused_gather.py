"""Two-GPU check of the fused all-gather (run under torchrun by tests/test_parallel.py or by hand):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/mp_fused_gather.py

Every rank encodes its own samples; the last GEMM of the tower stores its rows into every rank's gathered buffer
(NVLink peer stores). The result must equal, bit for bit, the NCCL all-gather of the ranks' local bf16 outputs.
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llm_quest_b200 import parallel  # noqa: E402
from llm_quest_b200.qwen.qwen3_5.qwen3_5_vision_model import Qwen3_5VisionModel  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = {"vision_emb_dim": 768, "vision_n_layers": 2, "vision_num_heads": 12, "vision_hidden_dim": 3072,
           "vision_rope_base": 10_000, "llm_d_in": 1024, "img_width": 128, "img_height": 128, "patch_size": 16,
           "in_channels": 3, "temporal_patch_size": 2, "spatial_merge_size": 2, "num_position_embeddings": 2304}
    torch.manual_seed(123)
    model = Qwen3_5VisionModel(cfg).eval().cuda()
    B, n_out = 3, (128 // 32) ** 2
    fg = parallel.FusedAllGather(rows_local=B * n_out, cols=1024)
    if rank == 0:
        print(f"multicast_ptr={'yes' if fg.multicast_ptr else 'no'} (hw support: {bool(getattr(fg.hdl, 'has_multicast_support', lambda *_: False))})", flush=True)
    ok = True
    for step in range(3):     # three steps: both slots are reused
        g = torch.Generator().manual_seed(100 * step + rank)
        x = torch.randn(B, 3, 2, 128, 128, generator=g).to(torch.bfloat16).cuda()
        with torch.inference_mode():
            local_out = model(x).to(torch.bfloat16)
            ref = parallel.all_gather_cat(local_out, 0)
            got = model(x, gather=fg)
        torch.cuda.synchronize()
        ok = ok and got.shape == ref.shape and torch.equal(got, ref)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("FUSED_GATHER_OK" if int(flag.item()) == 1 else "FUSED_GATHER_MISMATCH", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""CPU (gloo, world_size 2): the sample-sharding plumbing of the N>1 path.

Rank r handles samples [lo, hi) and the fused outputs are all-gathered; since no op of the path mixes
samples the gathered per-shard results must equal the single-process result. The
per-shard compute here is the CPU oracle (the CUDA kernels need a GPU); the collective plumbing —
bounds, dim-0 gather of embeddings, dim-1 gather of the [3, b, seq] position ids — is the product code.
"""

import os
from pathlib import Path
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from llm_quest_b200 import parallel
from oracle import fusion_oracle as FO

IMG = 248056


def test_shard_bounds_cover_and_balance():
    for n in (1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ids_np, table_np, vis_np, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ids = torch.from_numpy(ids_np)
        local_ids = parallel.shard_batch(ids)
        lo, hi = parallel.shard_bounds(ids.shape[0], rank, world)
        assert local_ids.shape[0] == hi - lo
        # every sample carries exactly n_vis placeholders, so shard-local scatter == global scatter
        n_vis = vis_np.shape[0] // ids.shape[0]
        local_vis = vis_np[lo * n_vis : hi * n_vis]
        fused = FO.fuse_embeddings(local_ids.numpy(), table_np, local_vis, image_token_id=IMG % table_np.shape[0])
        pid = FO.mrope_position_ids(local_ids.numpy(), [[1, 4, 4]], None, IMG % table_np.shape[0], 2)
        embs, pids = parallel.all_gather_fused(torch.from_numpy(fused.astype(np.int32)), torch.from_numpy(pid))
        if rank == 0:
            torch.save({"embs": embs, "pids": pids}, os.path.join(out_dir, "gathered.pt"))
        x = torch.full((2, 3), float(rank))
        g = parallel.all_gather_cat(x, dim=1)
        assert g.shape == (2, 3 * world) and torch.equal(g[:, :3], torch.zeros(2, 3))
    finally:
        dist.destroy_process_group()


def test_sharded_fuse_matches_global(tmp_path):
    world, b, seq, D, vocab = 2, 4, 30, 8, 50
    tok = IMG % vocab
    rng = np.random.default_rng(0)
    ids = rng.integers(0, vocab - 1, size=(b, seq))
    ids[ids == tok] = 0
    for s in range(b):
        ids[s, 3 + s : 7 + s] = tok  # 4 placeholders = one [1,4,4] feed after 2x2 merge
    table = rng.integers(0, 65535, size=(vocab, D)).astype(np.uint16)
    vis = rng.integers(0, 65535, size=(b * 4, D)).astype(np.uint16)
    mp.spawn(_worker, args=(world, _free_port(), ids, table, vis, str(tmp_path)), nprocs=world, join=True)
    got = torch.load(tmp_path / "gathered.pt")
    exp_embs = FO.fuse_embeddings(ids, table, vis, image_token_id=tok).astype(np.int32)
    exp_pid = FO.mrope_position_ids(ids, [[1, 4, 4]], None, tok, 2)
    assert np.array_equal(got["embs"].numpy(), exp_embs)
    assert np.array_equal(got["pids"].numpy(), exp_pid)


def test_streamed_encoder_rejects_cpu_model():
    from llm_quest_b200.pipeline import StreamedEncoder

    with pytest.raises(RuntimeError):
        StreamedEncoder(torch.nn.Linear(2, 2))


@pytest.mark.gpu
def test_fused_all_gather_two_gpus():
    """The all-gather fused into the last GEMM's epilogue (peer stores over NVLink) == NCCL all-gather, bit for bit."""
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one box")
    root = Path(__file__).resolve().parent.parent
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29577", str(root / "tests" / "mp_fused_gather.py")],
                       capture_output=True, text=True, timeout=600)
    assert "FUSED_GATHER_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]

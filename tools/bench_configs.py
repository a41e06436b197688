#!/usr/bin/env python
"""Per-config measurements for the BASELINE.json configs that are NOT the bench.py headline line.

    python tools/bench_configs.py [cfg3] [cfg3native] [cfg4] [cfg5] [mrope] > gpurun_out/configs.jsonl

One JSON line per config: ms/step (CUDA events, 3 warm-up + K timed steps, inputs resident in HBM and
larger than L2), samples/s, algorithmic TFLOP/s, and the per-kernel-family breakdown from
_lib.KernelTimer (tensor kernels in TFLOP/s, HBM-bound kernels in GB/s of algorithmic bytes next to the
copy peak). Synthetic inputs and seeds as SURVEY.md §8(d). Copy the output into profiles/.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from llm_quest_b200 import _lib  # noqa: E402
from llm_quest_b200.qwen.qwen3_5.qwen3_5_vision_model import Qwen3_5VisionModel  # noqa: E402
from llm_quest_b200.qwen.qwen3_5.qwen3_5_vlm_model import Qwen3_5VLM  # noqa: E402

PEAKS = bench.measured_peaks()
DEV = "cuda"


def timed(fn, steps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def families(fn, reps=2):
    with _lib.KernelTimer() as kt:
        for _ in range(reps):
            fn()
    torch.cuda.synchronize()
    fam = kt.summary()
    tot = sum(d["ms_total"] for d in fam.values())
    out = {}
    for k, d in sorted(fam.items(), key=lambda kv: -kv[1]["ms_total"]):
        e = {"launches": d["launches"] // reps, "ms_per_step": round(d["ms_total"] / reps, 4), "share": round(d["ms_total"] / tot, 3)}
        if d["flops"]:
            e["tflops"] = round(d["flops"] / (d["ms_total"] / 1e3) / 1e12, 1)
            e["frac_of_bf16_peak"] = round(e["tflops"] / PEAKS["bf16_tflops"], 3)
        if d["bytes"]:
            e["gbs"] = round(d["bytes"] / (d["ms_total"] / 1e3) / 1e9, 1)
            e["frac_of_hbm_peak"] = round(e["gbs"] / PEAKS["hbm_gbs"], 3)
        out[k] = e
    return out


def tower(px, npos=2304):
    cfg = bench.qwen_cfg(px)
    cfg["num_position_embeddings"] = npos
    torch.manual_seed(123)
    return Qwen3_5VisionModel(cfg).eval().to(DEV), cfg


def emit(name, workload, ms, samples, flops, fam, extra=None):
    line = {"config": name, "workload": workload, "ms_per_step": round(ms, 3), "samples_per_s": round(samples / (ms / 1e3), 1),
            "step_tflops": round(flops / (ms / 1e3) / 1e12, 1), "peaks": PEAKS, "kernels": fam}
    line.update(extra or {})
    print(json.dumps(line), flush=True)


def vlm_inputs(b, n_img, n_vis_per_img, text_len, rng):
    """c0,I,c1,I,...: text split as evenly as cfg-3 does (410,410,410,410,408 for 2048/4 images)."""
    chunks = [text_len // (n_img + 1) + (1 if i < text_len % (n_img + 1) else 0) for i in range(n_img + 1)]
    rows = []
    for _ in range(b):
        parts = []
        for i, c in enumerate(chunks):
            parts.append(torch.randint(0, 1000, (c,), generator=rng))
            if i < n_img:
                parts.append(torch.full((n_vis_per_img,), bench.IMG_TOKEN, dtype=torch.int64))
        rows.append(torch.cat(parts))
    return torch.stack(rows)


def cfg3(native):
    b, n_img = 32, 4
    torch.manual_seed(123)
    cfg = bench.qwen_cfg(448)
    vlm = Qwen3_5VLM(cfg).eval().to(DEV)
    rng = torch.Generator().manual_seed(4321)
    ids = vlm_inputs(b, 1 if native else n_img, 784 if native else 196, 2048, rng).to(DEV)
    gp = torch.Generator().manual_seed(1234)
    if native:   # forward()-native: one 8-frame clip per sample, cross-frame attention, S=3136
        pixels = torch.randn(b, 3, 8, 448, 448, generator=gp).to(torch.bfloat16).to(DEV)
        feeds, S, nsamp = None, 3136, b
    else:        # 128 independent images, 4 per sample
        pixels = torch.randn(b * n_img, 3, 2, 448, 448, generator=gp).to(torch.bfloat16).to(DEV)
        feeds, S, nsamp = torch.tensor([[1, 28, 28]] * n_img), 784, b * n_img
    fn = lambda: vlm.encode_and_fuse(ids, pixels, feeds, check=False)
    with torch.inference_mode():
        embs, pid, mask = fn()
        ms = timed(fn, 5)
        fam = families(fn)
    emit("cfg3native" if native else "cfg3",
         f"Qwen3.5 early-fusion prefill front end: batch {b}, {'one 8-frame 448x448 clip' if native else '4 images 448x448'} per sample + 2048 text tokens, seq {ids.shape[1]}",
         ms, b, bench.tower_flops(S) * nsamp, fam,
         {"images_per_s": round(nsamp * (8 if native else 1) / (2 if native else 1) / (ms / 1e3), 1) if native else round(nsamp / (ms / 1e3), 1),
          "max_position_id": int(pid.max().item()), "placeholders": int(mask.sum().item()), "inputs_embs_shape": list(embs.shape)})


def cfg4():
    B, T, px = 16, 16, 448
    model, _ = tower(px)
    x = torch.randn(B, 3, T, px, px, generator=torch.Generator().manual_seed(1234)).to(torch.bfloat16).to(DEV)
    S = (T // 2) * (px // 16) ** 2
    fn = lambda: model(x)
    with torch.inference_mode():
        ms = timed(fn, 5)
        fam = families(fn)
    emit("cfg4", f"Qwen3.5 video path: batch {B}, {T} frames 448x448 -> T'=8, S={S}", ms, B, bench.tower_flops(S) * B, fam,
         {"frames_per_s": round(B * T / (ms / 1e3), 1)})


def cfg5():
    for px in (224, 448, 672, 896, 1120, 1344):
        B = max(8, int(round(64 * (448 / px) ** 2 / 8)) * 8)
        model, _ = tower(px, npos=7056)
        x = torch.randn(B, 3, 2, px, px, generator=torch.Generator().manual_seed(1234)).to(torch.bfloat16).to(DEV)
        S = (px // 16) ** 2
        fn = lambda: model(x)
        with torch.inference_mode():
            ms = timed(fn, 5)
            fam = families(fn)
        emit(f"cfg5_{px}", f"high-res sweep: {px}x{px}, T=2, batch {B}, S={S}, num_position_embeddings=7056", ms, B,
             bench.tower_flops(S) * B, fam)
        del model, x
        torch.cuda.empty_cache()


def mrope():
    """cfg-3 consumer-side check: zero-centred RMSNorm + MRoPE-I apply on q [32,8,2832,256], k [32,2,2832,256]."""
    b, seq = 32, 2832
    g = torch.Generator().manual_seed(99)
    q = torch.randn(b, 8, seq, 256, generator=g).to(torch.bfloat16).to(DEV)
    k = torch.randn(b, 2, seq, 256, generator=g).to(torch.bfloat16).to(DEV)
    inv = 1.0 / (1e7 ** (torch.arange(0, 32, dtype=torch.float32) / 32))
    ang = torch.arange(8192, dtype=torch.float32)[:, None] * inv[None]
    cos, sin = torch.cat([ang.cos()] * 2, -1).to(DEV), torch.cat([ang.sin()] * 2, -1).to(DEV)
    pid = torch.randint(0, 2104, (3, b, seq), generator=g).to(DEV)
    wq = torch.zeros(256, device=DEV)
    fn = lambda: (_lib.mrope_apply(q, cos, sin, pid, (11, 11, 10), norm_weight=wq), _lib.mrope_apply(k, cos, sin, pid, (11, 11, 10), norm_weight=wq))
    ms = timed(fn, 10)
    fam = families(fn)
    emit("mrope_apply", "q [32,8,2832,256] + k [32,2,2832,256] bf16: zero-centred RMSNorm + MRoPE-I (one layer's worth)", ms, b, 0.0, fam)


def cfg1():
    """cfg-1: Part-1 ViT-B/16 classifier, 224x224, batch 8 (the reference's own CPU-runnable case) + batch 256."""
    from llm_quest_b200.multimodal.vision_transformer.vit_model import ViTModel

    cfg = {"img_width": 224, "img_height": 224, "patch_size": 16, "num_channels": 3, "emb_dim": 768, "n_layers": 12,
           "n_heads": 12, "drop_rate": 0.1, "qkv_bias": True, "num_classes": 100}
    torch.manual_seed(123)
    m = ViTModel(cfg).eval().to(DEV)
    flops_img = 2 * 196 * 768 * 768 + 12 * (2 * 197 * 768 * 2304 + 4 * 197 * 197 * 768 + 2 * 197 * 768 * 768 + 4 * 197 * 768 * 3072) + 2 * 768 * 100
    for B in (8, 256):
        x = torch.randn(B, 3, 224, 224, generator=torch.Generator().manual_seed(1234)).to(DEV)
        fn = lambda: m(x)
        with torch.inference_mode():
            ms = timed(fn, 10)
            fam = families(fn)
        emit(f"cfg1_b{B}", f"Part-1 ViT-B/16 classifier forward, 224x224, batch {B}, fp32 parameters / bf16 operands", ms, B,
             flops_img * B, fam)


def latency():
    """Small-batch latency of the Qwen tower at 448x448: eager launches vs CUDA-graph replay (pipeline.GraphedEncoder)."""
    from llm_quest_b200.pipeline import GraphedEncoder

    model, _ = tower(448)
    for B in (1, 4):
        x = torch.randn(B, 3, 2, 448, 448, generator=torch.Generator().manual_seed(1234)).to(torch.bfloat16).to(DEV)
        with torch.inference_mode():
            ms_eager = timed(lambda: model(x), 20)
        ge = GraphedEncoder(model, x)
        ms_graph = timed(lambda: ge(x), 20)
        print(json.dumps({"config": f"latency_b{B}", "workload": f"Qwen3-ViT tower 448x448, batch {B}: ms per forward",
                          "ms_eager": round(ms_eager, 3), "ms_cuda_graph": round(ms_graph, 3),
                          "images_per_s_graph": round(B / (ms_graph / 1e3), 1)}), flush=True)


def textattn():
    """cfg-3 consumer: MRoPEGatedAttention prefill on the fused embeddings [32, 2832, 1024] with the MRoPE-I ids."""
    from llm_quest_b200.common.rope import RoPE
    from llm_quest_b200.qwen.qwen3_5.qwen3_5_text_model import MRoPEGatedAttention

    cfg = {"emb_dim": 1024, "n_heads": 8, "num_kv_groups": 2, "head_dim": 256, "dtype": torch.bfloat16, "p_dropout": 0.0,
           "training": False, "mrope_section": [11, 11, 10]}
    torch.manual_seed(123)
    att = MRoPEGatedAttention(cfg, layer_idx=3).eval().to(DEV)
    b, seq = 32, 2832
    g = torch.Generator().manual_seed(99)
    x = torch.randn(b, seq, 1024, generator=g).to(torch.bfloat16).to(DEV)
    pid = torch.randint(0, 2104, (3, b, seq), generator=g).to(DEV)
    cos, sin = RoPE.compute_angles(10_000_000, 256, 8192, rotation_factor=0.25)
    cos, sin = cos.to(DEV), sin.to(DEV)
    fn = lambda: att(x, None, cos, sin, position_ids=pid)
    with torch.inference_mode():
        ms = timed(fn, 10)
        fam = families(fn)
    M = b * seq
    flops = 2.0 * M * 1024 * 5120 + 2.0 * M * 2048 * 1024 + 4.0 * b * 8 * seq * seq * 256 * 0.5
    emit("text_attention_prefill", "MRoPEGatedAttention prefill, one layer: batch 32 x seq 2832, 8 q / 2 kv heads x 256 "
         "(fused q|gate|k|v GEMM, in-place RMSNorm+MRoPE-I, causal GQA attention with sigmoid gate, out_proj)", ms, b, flops, fam,
         {"tokens_per_s": round(M / (ms / 1e3), 1)})


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg3", "cfg3native", "cfg4", "cfg5", "mrope", "textattn", "cfg1", "latency"]
    _lib.lib()
    for w in which:
        {"cfg3": lambda: cfg3(False), "cfg3native": lambda: cfg3(True), "cfg4": cfg4, "cfg5": cfg5, "mrope": mrope, "textattn": textattn, "cfg1": cfg1, "latency": latency}[w]()
        torch.cuda.empty_cache()

#!/usr/bin/env python
"""bench.py — images/sec of the vision encode (+ fuse) hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--device cpu|cuda]
                    [--workload cfg1|cfg2|cfg3|cfg4|cfg5:<px>] [--no-graph] [--no-cpu] [--no-eager] [--no-u8]

Workloads = BASELINE.json `configs`; the default and the configuration the metric is quoted on is cfg2:
  cfg1       Part-1 ViT-B/16 classifier forward, 224x224, batch 8 per GPU, fp32 parameters
  cfg2       Qwen3.5 Qwen3-ViT tower + 2x2 spatial-merge adapter, 448x448 (T=2), batch 64 per GPU
  cfg3       early-fusion prefill front end: 32 samples x (4 images 448x448 + 2048 text tokens), seq 2832: tower on the
             128 images + embedding gather + scatter of the merged rows + MRoPE-I position ids
  cfg4       video path: 16 clips x 16 frames 448x448 (T'=8, S=6272)
  cfg5:<px>  high-res sweep point, px in {224,...,1344}, T=2, batch 64*(448/px)^2 (multiple of 8, at least 8) per GPU
Random-init weights (seed 123), synthetic randn pixels (seed 1234), token ids (seed 4321). One step = one pass of the
whole batch. With N GPUs every rank runs its own batch (sample sharding, weak scaling); for the tower workloads the
merged embeddings are all-gathered inside the step, fused into the last GEMM's epilogue (parallel.FusedAllGather).

One JSON line on stdout (rank 0):
  value       images/s with the fp32 pixel batch already resident in HBM (CUDA events, max over ranks); the step is
              replayed from a CUDA graph (pipeline.GraphedEncoder) unless --no-graph
  e2e         same metric through the public streaming API with HOST (pinned, fp32) pixels: H2D copy of the batch and
              D2H read of the result inside the timed region, every step
  roofline    the dominant kernel family (tcgen05 GEMM): algorithmic FLOPs / measured per-launch time against the
              measured cuBLAS peaks (burst and sustained) and the nominal 2.25 PFLOP/s
  gpu_eager_baseline  the UNMODIFIED reference modules (baseline/_ref) `.to("cuda", bfloat16)` on the same GPU, same
              seeds and shapes, CUDA-event timed in this process: cuBLASLt GEMMs + SDPA + ATen elementwise — the
              "existing Blackwell kernels" bar
  cpu_baseline  the reference itself (kind "reference"; the oracle port only if baseline/_ref is missing) on this
              box's host cores, bounded sample
`--impl reference` runs the reference alone (rank 0 only): on the host cores by default, `--device cuda` for the eager
bf16 GPU arm; no repo .so is loaded in that arm.
"""

from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

IMG_TOKEN = 248056
METRIC = "images/sec (vision encode+fuse)"
NOMINAL_BF16_TFLOPS = 2250.0


def qwen_cfg(px=448, npos=2304):
    return {
        "vision_emb_dim": 768, "vision_n_layers": 12, "vision_num_heads": 12, "vision_hidden_dim": 3072,
        "vision_rope_base": 10_000, "llm_d_in": 1024, "img_width": px, "img_height": px, "patch_size": 16,
        "in_channels": 3, "temporal_patch_size": 2, "spatial_merge_size": 2, "num_position_embeddings": npos,
        "image_token_id": IMG_TOKEN, "vocab_size": 248_320, "emb_dim": 1024, "dtype": torch.bfloat16,
    }


VIT_CFG = {"img_width": 224, "img_height": 224, "patch_size": 16, "num_channels": 3, "emb_dim": 768, "n_layers": 12,
           "n_heads": 12, "drop_rate": 0.1, "qkv_bias": True, "num_classes": 100}


def tower_flops(S: int) -> float:
    """Algorithmic FLOPs per sample: 2MNK per GEMM, 4*S^2*D per attention layer."""
    return (2 * S * 1536 * 768 + 12 * (2 * S * 768 * 2304 + 4 * S * S * 768 + 2 * S * 768 * 768 + 4 * S * 768 * 3072)
            + (S // 4) * (2 * 3072 * 3072 + 2 * 3072 * 1024))


VIT_FLOPS_PER_IMG = (2 * 196 * 768 * 768 + 12 * (2 * 197 * 768 * 2304 + 4 * 197 * 197 * 768 + 2 * 197 * 768 * 768
                                                 + 4 * 197 * 768 * 3072) + 2 * 768 * 100)


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"bf16_tflops_sustained": d.get("bf16_tflops_sustained", d.get("bf16_tflops")), "bf16_tflops_burst": d.get("bf16_tflops"),
                "hbm_gbs": d.get("hbm_gbs"), "sm_max_mhz": d.get("sm_max_mhz", 1965.0), "source": "measured"}
    return {"bf16_tflops_sustained": 1400.0, "bf16_tflops_burst": 1590.0, "hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "source": "fallback"}


def measured_traffic():
    """DRAM bytes per launch of the dominant kernel family from the newest committed ncu --set full capture."""
    def order(f):   # r02_v3 after r01_v14: numeric, not lexicographic
        return [int(n) for n in re.findall(r"\d+", f.name)]

    files = sorted((ROOT / "profiles").glob("*_traffic.json"), key=order)
    if not files:
        return None, None
    d = json.loads(files[-1].read_text())
    return d.get("gemm_family_avg_dram_bytes_per_launch"), files[-1].name


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in self.rows)]
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------
def vlm_input_ids(b, n_img, n_vis_per_img, text_len, rng):
    """c0,I,c1,I,...: text split as evenly as cfg-3 asks (410,410,410,410,408 for 2048 tokens / 4 images)."""
    chunks = [text_len // (n_img + 1) + (1 if i < text_len % (n_img + 1) else 0) for i in range(n_img + 1)]
    rows = []
    for _ in range(b):
        parts = []
        for i, c in enumerate(chunks):
            parts.append(torch.randint(0, 1000, (c,), generator=rng))
            if i < n_img:
                parts.append(torch.full((n_vis_per_img,), IMG_TOKEN, dtype=torch.int64))
        rows.append(torch.cat(parts))
    return torch.stack(rows)


class Workload:
    """One BASELINE.json configuration: shapes, synthetic inputs, our step, the reference's step, FLOPs."""

    key = ""
    gather = False           # N > 1: all-gather of the merged embeddings fused into the last GEMM

    def __init__(self, batch=None, cpu_batch=None):
        self.B = batch or self.default_batch
        self.cpu_B = cpu_batch or self.default_cpu_batch

    # --- inputs -------------------------------------------------------------------------------
    def host_inputs(self, rank=0, batch=None):
        raise NotImplementedError

    def images(self, batch=None):
        """image-equivalents per step per GPU (the metric's unit)."""
        return batch or self.B

    def flops(self, batch=None):
        raise NotImplementedError

    # --- ours ---------------------------------------------------------------------------------
    def build_ours(self, dev):
        raise NotImplementedError

    def ours_step(self, model, dev_in, gather=None):
        """One pass on device-resident inputs; returns the tensor(s) a caller reads back."""
        raise NotImplementedError

    # --- the unmodified reference (baseline/_ref) ----------------------------------------------
    def build_reference(self):
        raise NotImplementedError

    def reference_step(self, ref, inputs):
        raise NotImplementedError


class TowerWorkload(Workload):
    """Qwen3-ViT tower + merge adapter (cfg2 / cfg4 / cfg5)."""

    gather = True

    def __init__(self, key, px, T, default_batch, default_cpu_batch, npos=2304, **kw):
        self.key, self.px, self.T, self.npos = key, px, T, npos
        self.default_batch, self.default_cpu_batch = default_batch, default_cpu_batch
        super().__init__(**kw)
        self.S = (T // 2) * (px // 16) ** 2
        self.n_out = self.S // 4

    def describe(self, B):
        kind = {"cfg2": "Qwen3.5 Qwen3-ViT tower + spatial-merge adapter", "cfg4": "Qwen3.5 video path (tower + merge adapter)"}.get(
            self.key.split(":")[0], "high-res sweep point (tower + merge adapter)")
        return f"{self.key}: {kind}, {self.px}x{self.px}, T={self.T} (S={self.S} tokens/sample), batch {B} per GPU"

    def host_inputs(self, rank=0, batch=None):
        B = batch or self.B
        g = torch.Generator().manual_seed(1234 + rank)
        return {"pixels": torch.randn(B, 3, self.T, self.px, self.px, generator=g)}

    def images(self, batch=None):
        # a video sample counts as T/2 temporal patches = image-equivalents of the T=2 image format
        return (batch or self.B) * (self.T // 2)

    def flops(self, batch=None):
        return tower_flops(self.S) * (batch or self.B)

    def build_ours(self, dev):
        from llm_quest_b200.qwen.qwen3_5.qwen3_5_vision_model import Qwen3_5VisionModel

        torch.manual_seed(123)
        return Qwen3_5VisionModel(qwen_cfg(self.px, self.npos)).eval().to(dev)

    def ours_step(self, model, dev_in, gather=None):
        if gather is not None:
            return model(dev_in["pixels"], gather=gather)
        return model(dev_in["pixels"])

    def build_reference(self):
        from baseline import ref

        torch.manual_seed(123)
        return ref.qwen_vision_model({"img_width": self.px, "img_height": self.px, "num_position_embeddings": self.npos}).eval()

    def reference_step(self, ref_model, inputs):
        return ref_model(inputs["pixels"])


class FuseWorkload(Workload):
    """cfg3: tower on 128 images + embedding gather + masked scatter + MRoPE-I position ids (everything of
    Qwen3_5VLM.forward before the text model, qwen3_5_vlm_model.py:198-218)."""

    key = "cfg3"
    default_batch, default_cpu_batch = 32, 1
    n_img, text_len, px = 4, 2048, 448

    def describe(self, B):
        return (f"cfg3: Qwen3.5 early-fusion prefill front end, batch {B} per GPU x ({self.n_img} images 448x448 + {self.text_len} text tokens), "
                f"seq {self.text_len + self.n_img * 196}: tower + embedding gather + scatter + MRoPE-I position ids")

    def host_inputs(self, rank=0, batch=None):
        B = batch or self.B
        ids = vlm_input_ids(B, self.n_img, 196, self.text_len, torch.Generator().manual_seed(4321 + rank))
        px = torch.randn(B * self.n_img, 3, 2, self.px, self.px, generator=torch.Generator().manual_seed(1234 + rank))
        return {"ids": ids, "pixels": px}

    def images(self, batch=None):
        return (batch or self.B) * self.n_img

    def flops(self, batch=None):
        return tower_flops(784) * (batch or self.B) * self.n_img

    def build_ours(self, dev):
        from llm_quest_b200.qwen.qwen3_5.qwen3_5_vlm_model import EmbeddingOnlyLM, Qwen3_5VLM

        torch.manual_seed(123)
        cfg = qwen_cfg(self.px)
        return Qwen3_5VLM(cfg, language_model=EmbeddingOnlyLM(cfg)).eval().to(dev)

    def ours_step(self, model, dev_in, gather=None):
        feeds = torch.tensor([[1, 28, 28]] * self.n_img)
        embs, pid, _ = model.encode_and_fuse(dev_in["ids"], dev_in["pixels"], feeds, check=False)
        return embs, pid

    def build_reference(self):
        """The reference VLM with its own forward; the (out-of-scope) text model is cut off right at its call site:
        a 1-layer text model is built for its emb_dict and its forward is replaced by a pass-through, so that
        Qwen3_5VLM.forward returns (inputs_embs, position_ids) — lines 198-218 run unmodified."""
        from baseline import ref

        torch.manual_seed(123)
        vlm = ref.qwen_vlm({"img_width": self.px, "img_height": self.px, "n_layers": 1}).eval()
        vlm.language_model.forward = lambda inputs_embs=None, position_ids=None, attn_mask=None, **kw: (inputs_embs, position_ids)
        # forward() derives ONE feed from the pixel batch; cfg-3 has 4 independent images per sample
        feeds = torch.tensor([[1, 28, 28]] * self.n_img)
        vlm.get_feeds_3d_shape = lambda image_pixels: feeds
        return vlm

    def reference_step(self, ref_model, inputs):
        return ref_model(inputs["ids"], image_pixels=inputs["pixels"])


class ViTWorkload(Workload):
    """cfg1: Part-1 ViT-B/16 classifier forward."""

    key = "cfg1"
    default_batch, default_cpu_batch = 8, 8

    def describe(self, B):
        return f"cfg1: Part-1 ViT-B/16 classifier forward, 224x224, batch {B} per GPU, fp32 parameters"

    def host_inputs(self, rank=0, batch=None):
        return {"pixels": torch.randn(batch or self.B, 3, 224, 224, generator=torch.Generator().manual_seed(1234 + rank))}

    def flops(self, batch=None):
        return VIT_FLOPS_PER_IMG * (batch or self.B)

    def build_ours(self, dev):
        from llm_quest_b200.multimodal.vision_transformer.vit_model import ViTModel

        torch.manual_seed(123)
        return ViTModel(dict(VIT_CFG)).eval().to(dev)

    def ours_step(self, model, dev_in, gather=None):
        return model(dev_in["pixels"])

    def build_reference(self):
        from baseline import ref

        torch.manual_seed(123)
        return ref.vit_model().eval()

    def reference_step(self, ref_model, inputs):
        return ref_model(inputs["pixels"])


def make_workload(name: str, batch=None, cpu_batch=None) -> Workload:
    kw = {"batch": batch, "cpu_batch": cpu_batch}
    if name == "cfg1":
        return ViTWorkload(**kw)
    if name == "cfg2":
        return TowerWorkload("cfg2", 448, 2, 64, 8, **kw)
    if name == "cfg3":
        return FuseWorkload(**kw)
    if name == "cfg4":
        return TowerWorkload("cfg4", 448, 16, 16, 1, **kw)
    m = re.fullmatch(r"cfg5:(\d+)", name)
    if m:
        px = int(m.group(1))
        if px % 32 or not 32 <= px <= 1344:
            raise SystemExit("cfg5:<px>: px must be a multiple of 32 (patch 16 x merge 2) up to 1344")
        B = max(8, int(round(64 * (448 / px) ** 2 / 8)) * 8)
        return TowerWorkload(name, px, 2, B, max(1, min(8, B // 8)), npos=7056, **kw)
    raise SystemExit(f"unknown workload {name!r} (cfg1, cfg2, cfg3, cfg4, cfg5:<px>)")


# ------------------------------------------------------------------------------------------------
# the reference arms
# ------------------------------------------------------------------------------------------------
def reference_cpu_run(wl: Workload, steps: int, warmup: int):
    """The reference's own modules on the host cores (fp32, inference_mode, all threads). Falls back to the oracle
    port (kind "port", cfg2-style towers only) when baseline/_ref is not there."""
    from baseline import ref

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = wl.cpu_B
    inputs = wl.host_inputs(0, B)
    if ref.available():
        kind = "reference"
        model = wl.build_reference()
        fn = lambda: wl.reference_step(model, inputs)
    else:
        if not isinstance(wl, TowerWorkload):
            raise SystemExit("baseline/_ref missing (run baseline/install_ref.sh) and the oracle port only covers the tower workloads")
        from oracle import vision_oracle as VO
        from llm_quest_b200.qwen.qwen3_5.qwen3_5_vision_model import Qwen3_5VisionModel

        kind = "port"
        cfg = qwen_cfg(wl.px, wl.npos)
        torch.manual_seed(123)
        sd = {k: v.detach() for k, v in Qwen3_5VisionModel(cfg).state_dict().items()}   # parameter containers only
        fn = lambda: VO.qwen_vision_forward(sd, cfg, inputs["pixels"])
    times = []
    with torch.inference_mode():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            fn()
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    total = sum(times)
    return {"img_per_s": wl.images(B) * len(times) / total, "ms_per_step": 1e3 * total / len(times), "cores": cores, "kind": kind,
            "sample": f"{wl.describe(B)} — bounded sample of {B} sample(s) per step, fp32, {len(times)} timed steps after {warmup} warm-up"}


def reference_cuda_run(wl: Workload, steps: int, warmup: int, dev, batch=None):
    """The unmodified reference modules `.to(cuda, bfloat16)` (PyTorch eager: cuBLASLt, SDPA, ATen) on the GPU."""
    B = batch or wl.B
    host = wl.host_inputs(0, B)
    model = wl.build_reference().to(device=dev, dtype=torch.bfloat16)
    inputs = {k: (v.to(dev, torch.bfloat16) if v.is_floating_point() else v.to(dev)) for k, v in host.items()}
    with torch.inference_mode():
        for _ in range(max(warmup, 2)):
            wl.reference_step(model, inputs)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = wl.reference_step(model, inputs)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del model, inputs, out
    torch.cuda.empty_cache()
    try:
        sdpa = [n for n, f in (("flash", torch.backends.cuda.flash_sdp_enabled), ("mem_efficient", torch.backends.cuda.mem_efficient_sdp_enabled),
                               ("cudnn", torch.backends.cuda.cudnn_sdp_enabled), ("math", torch.backends.cuda.math_sdp_enabled)) if f()]
    except Exception:  # noqa: BLE001
        sdpa = []
    return {"value": round(wl.images(B) / (ms / 1e3), 1), "unit": "images/s", "ms_per_step": round(ms, 3), "batch": B, "steps": steps,
            "step_tflops": round(wl.flops(B) / (ms / 1e3) / 1e12, 1), "dtype": "bf16",
            "what": "unmodified reference modules (baseline/_ref) .to('cuda', bfloat16), PyTorch eager, inputs resident, CUDA events",
            "sdpa_backends_enabled": sdpa}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = make_workload(args.workload, args.batch, args.cpu_batch)
    base = {"impl": "reference", "metric": METRIC, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "data": "synthetic", "gpu_launches": 0}
    if args.device == "cuda":
        if not torch.cuda.is_available():
            raise SystemExit("--device cuda needs a GPU")
        torch.cuda.set_device(0)
        r = reference_cuda_run(wl, max(1, args.steps), max(3, args.warmup), torch.device("cuda", 0))
        line = {**base, "value": r["value"], "ms_per_step": r["ms_per_step"], "dtype": "bf16", "device": "cuda",
                "config": {"workload": wl.describe(r["batch"]) + " (reference modules in PyTorch eager bf16 on the GPU)"},
                "gpu_eager_baseline": r,
                "e2e": {"value": r["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    else:
        r = reference_cpu_run(wl, max(1, args.steps), max(1, min(args.warmup, 1)))
        v = round(r["img_per_s"], 3)
        line = {**base, "value": v, "ms_per_step": round(r["ms_per_step"], 2), "dtype": "f32", "device": "cpu",
                "config": {"workload": wl.describe(wl.cpu_B) + " (CPU arm: bounded sample; no op mixes samples)"},
                "cpu_baseline": {"value": v, "unit": "images/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
                "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU path
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist

    from llm_quest_b200 import _lib, parallel
    from llm_quest_b200.pipeline import GraphedEncoder, StreamedEncoder

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the vision-encode-and-fuse kernels have no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()

    # host side of the end-to-end path: this rank's pinned buffers should live on the NUMA node next to its GPU
    from llm_quest_b200.pipeline import bind_host_thread_to_gpu_numa

    numa = bind_host_thread_to_gpu_numa(local) if world > 1 else "single rank: not bound"
    wl = make_workload(args.workload, args.batch, args.cpu_batch)
    B = wl.B
    model = wl.build_ours(dev)
    host = {k: v.pin_memory() for k, v in wl.host_inputs(rank).items()}      # fp32 pixels / int64 ids, as the callers hold them
    dev_in = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    in_bytes = sum(v.numel() * v.element_size() for v in host.values())

    # N > 1 (tower workloads): the all-gather is FUSED into the tower's last GEMM — its epilogue stores this rank's rows
    # into every rank's gathered buffer over NVLink (parallel.FusedAllGather) followed by a signal-pad barrier.
    fused, overlap, gather_kind = None, None, "none"
    if world > 1 and wl.gather:
        try:
            fused = parallel.FusedAllGather(rows_local=B * wl.n_out, cols=1024, slots=3)
            gather_kind = ("all-gather fused into the last GEMM (" + ("NVSwitch multimem stores" if fused.multicast_ptr else "NVLink peer stores")
                           + " + signal barrier)")
        except Exception as e:  # noqa: BLE001
            print(f"[bench] fused all-gather unavailable ({type(e).__name__}: {e}); using NCCL", file=sys.stderr, flush=True)
            overlap = parallel.OverlappedAllGather(dev)
            gather_kind = "NCCL all-gather of merged embeddings on a side stream"

    def step_eager(inputs=dev_in):
        if fused is not None:
            fused.set_next_slot(0)       # slot 0: the resident step; slots 1, 2: the two ring slots of the e2e pipeline
        out = wl.ours_step(model, inputs, gather=fused)
        if overlap is not None:
            out, _ = overlap.submit(out.to(torch.bfloat16))
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    gather_check = None
    with torch.inference_mode():
        for _ in range(2):
            step_eager()
        torch.cuda.synchronize()
        if fused is not None:
            # SCALE itself proves the fused collective: every rank's gathered buffer must equal an NCCL all-gather of the
            # local rows (one step, outside the timed region), bit for bit.
            full = step_eager()
            mine = full[rank * B:(rank + 1) * B].contiguous()
            ref_full = parallel.all_gather_cat(mine, 0)
            ok = torch.tensor([int(torch.equal(full, ref_full))], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            gather_check = bool(ok.item())
            if not gather_check:
                raise SystemExit("bench.py: the fused all-gather differs from the NCCL all-gather")

        # launches of one step (counted on an eager pass; the graph replays exactly these kernel nodes)
        _lib.reset_launch_count()
        step_eager()
        torch.cuda.synchronize()
        launches_per_step = _lib.launch_count()

        # the timed step: a CUDA-graph replay of the whole pass (no Python between the launches)
        step, graphed = step_eager, False
        if not args.no_graph and overlap is None:
            try:
                ge = GraphedEncoder(lambda: step_eager(), None)
                step, graphed = ge.replay, True
            except Exception as e:  # noqa: BLE001
                print(f"[bench] CUDA-graph capture failed ({type(e).__name__}: {e}); timing eager launches", file=sys.stderr, flush=True)
                torch.cuda.synchronize()
        for _ in range(max(3, args.warmup)):
            step()
        with ClockSampler(local) as clocks:
            ms_total = timed(step, args.steps)

        # per-kernel timing for the roofline: event pairs around every launch, recorded INSIDE a second CUDA graph of the
        # step (external events), so every kernel is timed in the back-to-back regime of the timed region — same clocks,
        # same cache state; an eager pass with event pairs (idle gaps between launches, boost clocks) is the fallback
        fam, fam_reps, fam_mode = None, 2, "eager launches with event pairs"
        if graphed:
            try:
                with _lib.KernelTimer(external=True) as kt:
                    gt = GraphedEncoder(lambda: wl.ours_step(model, dev_in), None, warmup=0)
                for _ in range(max(3, args.steps)):      # as many replays as the timed region: the same clock regime
                    gt.replay()
                torch.cuda.synchronize()
                fam, fam_reps, fam_mode = kt.summary(), 1, "event pairs inside a CUDA-graph replay of the step"
                del gt
            except Exception as e:  # noqa: BLE001
                print(f"[bench] in-graph kernel timing unavailable ({type(e).__name__}: {e}); eager event pairs", file=sys.stderr, flush=True)
                torch.cuda.synchronize()
        if fam is None:
            with _lib.KernelTimer() as kt:
                for _ in range(2):
                    wl.ours_step(model, dev_in)
            torch.cuda.synchronize()
            fam = kt.summary()

        # end to end through the public streaming API: pinned host fp32 pixels (+ ids) in, result read back to pinned host
        # memory every step. N > 1: each rank reads back its own shard, the gathered batch stays in HBM for its consumer.
        def e2e_model(inputs):
            out = wl.ours_step(model, inputs, gather=fused)
            if fused is not None:
                return out[rank * B:(rank + 1) * B]
            if overlap is not None:
                local_rows = out.to(torch.bfloat16)
                overlap.submit(local_rows)
                return local_rows
            return out

        enc = StreamedEncoder(e2e_model, depth=2, device=dev, graph=not args.no_graph and overlap is None,
                              before_slot=(lambda i: fused.set_next_slot(1 + i)) if fused is not None else None)
        sink = [0.0]

        def run_e2e(steps):
            for _ in range(steps):
                enc.submit(host)
                for out in enc.ready():
                    sink[0] += float(out[0].flatten()[0])   # touch the host result
            for out in enc.drain():
                sink[0] += float(out[0].flatten()[0])

        run_e2e(3)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_e2e(args.steps)
        torch.cuda.synchronize()
        e1.record()
        barrier()
        t_e2e = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        ms_e2e = float(t_e2e.item())
        out_bytes = enc.last_out_bytes

        # Informational: the same end-to-end loop fed the way the reference's image path really starts — uint8
        # frames (qwen3_5_generate_multimodal.py:40-46 turns one image into fp32 pixels and repeats it on the temporal axis) —
        # with normalisation and the temporal repeat done by vf_preprocess_u8 on the device: 1/8 of the upload.
        e2e_u8 = None
        if isinstance(wl, TowerWorkload) and wl.T == 2 and not args.no_u8:
            from llm_quest_b200.qwen.qwen3_5.preprocess import pixels_from_uint8

            enc = None
            g8 = torch.Generator().manual_seed(4321 + rank)
            host_u8 = torch.randint(0, 256, (B, wl.px, wl.px, 3), dtype=torch.uint8, generator=g8).pin_memory()
            mean, std = (0.5, 0.5, 0.5), (0.5, 0.5, 0.5)

            def u8_model(frames):
                out = model(pixels_from_uint8(frames, mean, std, 2, torch.float32), gather=fused) if fused is not None else \
                    model(pixels_from_uint8(frames, mean, std, 2, torch.float32))
                return out[rank * B:(rank + 1) * B] if fused is not None else out

            enc8 = StreamedEncoder(u8_model, depth=2, device=dev, graph=not args.no_graph,
                                   before_slot=(lambda i: fused.set_next_slot(1 + i)) if fused is not None else None)

            def run_u8(steps):
                for _ in range(steps):
                    enc8.submit(host_u8)
                    for out in enc8.ready():
                        sink[0] += float(out[0].flatten()[0])
                for out in enc8.drain():
                    sink[0] += float(out[0].flatten()[0])

            run_u8(3)
            barrier()
            e0.record()
            run_u8(args.steps)
            torch.cuda.synchronize()
            e1.record()
            barrier()
            t8 = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(t8, op=dist.ReduceOp.MAX)
            e2e_u8 = {"value": round(world * wl.images() * args.steps / (float(t8.item()) / 1e3), 1), "unit": "images/s",
                      "ms_per_step": round(float(t8.item()) / args.steps, 3), "h2d_bytes_per_step": world * host_u8.numel(),
                      "d2h_bytes_per_step": world * enc8.last_out_bytes,
                      "note": "informational: uint8 frames [B, H, W, 3] in, normalise + temporal repeat on the device (vf_preprocess_u8); "
                              "same tower work per image, not the metric's input format"}
            enc8 = None

    ms_step = ms_total / args.steps
    imgs = wl.images()
    value = world * imgs * args.steps / (ms_total / 1e3)
    e2e_value = world * imgs * args.steps / (ms_e2e / 1e3)
    peaks = measured_peaks()

    line = None
    if rank == 0:
        is_gemm = lambda k: k.startswith("gemm_") or k == "patch_embed"
        gemm_ms = sum(d["ms_total"] for k, d in fam.items() if is_gemm(k))
        gemm_fl = sum(d["flops"] for k, d in fam.items() if is_gemm(k))
        gemm_n = sum(d["launches"] for k, d in fam.items() if is_gemm(k))
        all_ms = sum(d["ms_total"] for d in fam.values())
        achieved = gemm_fl / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
        breakdown = {}
        for k, d in sorted(fam.items(), key=lambda kv: -kv[1]["ms_total"]):
            e = {"launches": d["launches"] // fam_reps, "ms_per_step": round(d["ms_total"] / fam_reps, 3), "share": round(d["ms_total"] / all_ms, 3)}
            if d["flops"]:
                e["tflops"] = round(d["flops"] / (d["ms_total"] / 1e3) / 1e12, 1)
            if d["bytes"]:
                e["gbs"] = round(d["bytes"] / (d["ms_total"] / 1e3) / 1e9, 1)
            breakdown[k] = e
        ta_ms = sum(d["ms_total"] for k, d in fam.items() if is_gemm(k) or k == "attention")
        ta_fl = sum(d["flops"] for k, d in fam.items() if is_gemm(k) or k == "attention")
        step_tflops = wl.flops() / (ms_step / 1e3) / 1e12
        ck = clocks.summary()
        # the denominator that matches the clock regime the timed region actually ran in
        burst_regime = bool(ck["sm_mhz"]) and ck["sm_mhz"] >= 0.9 * (ck["sm_max_mhz"] or peaks["sm_max_mhz"])
        peak = peaks["bf16_tflops_burst"] if burst_regime else peaks["bf16_tflops_sustained"]
        traffic, traffic_src = measured_traffic()
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": round(ms_step, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": wl.describe(B), "global_batch": world * B, "images_per_step": world * imgs,
                       "l2_policy": f"inputs_larger_than_l2 ({in_bytes >> 20} MB of inputs, > 1 GB of activations per step)" if in_bytes > (126 << 20)
                       else f"activations_larger_than_l2 (inputs {in_bytes >> 20} MB; every layer streams its activations through HBM)",
                       "timed_step": "CUDA-graph replay of the whole pass" if graphed else "eager launches from Python",
                       "parallelism": f"sample-sharded x{world}" + (f" + {gather_kind}" if world > 1 and wl.gather else "")},
            "step_tflops": round(step_tflops, 1),
            "step_frac_of_peak": {"burst": round(step_tflops / peaks["bf16_tflops_burst"], 3), "sustained": round(step_tflops / peaks["bf16_tflops_sustained"], 3),
                                  "nominal": round(step_tflops / NOMINAL_BF16_TFLOPS, 3)},
            "kernel_time_share_of_step": round(all_ms / fam_reps / ms_step, 3), "kernel_timing": fam_mode,
            "roofline": {"bound": "tensor", "kernel": "vf::gemm_kernel<EPI,BN> (tcgen05, all epilogues; incl. patch-embed gather GEMM)",
                         "achieved": round(achieved, 1), "peak": peak, "unit": "TFLOP/s", "frac": round(achieved / peak, 3),
                         "peak_regime": ("burst" if burst_regime else "sustained") + f" cuBLAS bf16 ({peaks['source']}); timed region ran at a median of {ck['sm_mhz']} MHz",
                         "frac_burst": round(achieved / peaks["bf16_tflops_burst"], 3), "frac_sustained": round(achieved / peaks["bf16_tflops_sustained"], 3),
                         "frac_nominal": round(achieved / NOMINAL_BF16_TFLOPS, 3),
                         "traffic": traffic,
                         "traffic_unit": "DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, avg over the family)",
                         "traffic_source": traffic_src, "launches_per_step": gemm_n // fam_reps, "avg_launch_ms": round(gemm_ms / max(gemm_n, 1), 4),
                         "share_of_step": round(gemm_ms / all_ms, 3),
                         "gemm_plus_attention_tflops": round(ta_fl / (ta_ms / 1e3) / 1e12, 1) if ta_ms else None},
            "kernels": breakdown,
            "e2e": {"value": round(e2e_value, 1), "unit": "images/s", "ms_per_step": round(ms_e2e / args.steps, 3),
                    "h2d_bytes_per_step": world * in_bytes, "d2h_bytes_per_step": world * out_bytes,
                    "bytes_note": "whole job (all ranks); N>1: each rank reads back its own shard, the all-gathered batch stays in HBM",
                    "host_numa": numa,
                    "api": "llm_quest_b200.pipeline.StreamedEncoder: pinned-host fp32 pixels (+ int64 ids) in, result read back to pinned "
                           "host memory every step; upload / compute / download on 3 streams"},
            "gpu_launches": int(launches_per_step) * args.steps,
            "launches_per_step": int(launches_per_step),
            "clocks": ck,
        }
        if isinstance(wl, TowerWorkload):
            # a video sample counts as one sample and also as its T raw frames — report both next to the
            # image-equivalents (T/2 temporal patches) `value` is quoted in
            sps = world * B * args.steps / (ms_total / 1e3)
            line["samples_per_s"], line["raw_frames_per_s"] = round(sps, 1), round(sps * wl.T, 1)
        if e2e_u8 is not None:
            line["e2e_from_uint8_frames"] = e2e_u8
        if gather_check is not None:
            line["fused_gather_equals_nccl"] = gather_check
    # the baselines run on rank 0 at N = 1 only
    if rank == 0 and world == 1:
        enc = None
        torch.cuda.empty_cache()
        from baseline import ref

        if not args.no_eager and ref.available():
            try:
                line["gpu_eager_baseline"] = reference_cuda_run(wl, min(args.steps, 10), 3, dev)
                line["gpu_eager_baseline"]["ours_over_eager"] = round(value / line["gpu_eager_baseline"]["value"], 3)
            except Exception as e:  # noqa: BLE001
                line["gpu_eager_baseline"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
        if not args.no_cpu:
            r = reference_cpu_run(wl, 2, 1)
            line["cpu_baseline"] = {"value": round(r["img_per_s"], 3), "unit": "images/s", "cores": r["cores"], "kind": r["kind"],
                                    "sample": r["sample"]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--device", default="cpu", choices=["cpu", "cuda"], help="--impl reference: host cores (default) or eager bf16 on the GPU")
    ap.add_argument("--workload", default="cfg2", help="cfg1 | cfg2 | cfg3 | cfg4 | cfg5:<px>")
    ap.add_argument("--batch", type=int, default=None, help="samples per GPU per step (default: the workload's)")
    ap.add_argument("--cpu-batch", type=int, default=None, help="samples per step of the CPU baseline sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-eager", action="store_true", help="skip the gpu_eager_baseline leg")
    ap.add_argument("--no-u8", action="store_true", help="skip the informational uint8-frames end-to-end leg")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of a CUDA-graph replay")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

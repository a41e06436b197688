#!/usr/bin/env python
"""bench.py — images/sec of the vision encode (+ fuse) hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3]

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): Qwen3.5 Qwen3-ViT
tower + 2x2 spatial-merge adapter, 448x448 images (T=2 duplicated-frame format), batch 64 per GPU,
bf16 operands, random-init weights (seed 123), synthetic randn pixels (seed 1234). One step = one
forward of the whole batch. With N GPUs every rank encodes its own 64 samples (sample sharding,
weak scaling) and the merged embeddings are all-gathered over NCCL inside the step.

One JSON line on stdout (rank 0):
  value      images/s with the pixel batch already resident in HBM (CUDA events, max over ranks)
  e2e        same metric through the public nn.Module call with HOST (pinned) pixels: H2D copy of the
             batch and D2H read of the merged embeddings inside the timed region, every step
  roofline   the dominant kernel family (tcgen05 GEMM) — algorithmic FLOPs / measured per-launch time
             vs MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (oracle/vision_oracle.py, a port of the reference on the same ATen
             ops) timed on this box's host cores on a bounded sample
`--impl reference` times that CPU path alone (rank 0 only) and prints the same line shape.
"""

from __future__ import annotations

import argparse
import json
import re
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

IMG_TOKEN = 248056


def qwen_cfg(px=448):
    return {
        "vision_emb_dim": 768, "vision_n_layers": 12, "vision_num_heads": 12, "vision_hidden_dim": 3072,
        "vision_rope_base": 10_000, "llm_d_in": 1024, "img_width": px, "img_height": px, "patch_size": 16,
        "in_channels": 3, "temporal_patch_size": 2, "spatial_merge_size": 2, "num_position_embeddings": 2304,
        "image_token_id": IMG_TOKEN, "vocab_size": 248_320, "emb_dim": 1024, "dtype": torch.bfloat16,
    }


def tower_flops(S: int) -> float:
    """Algorithmic FLOPs per sample: 2MNK per GEMM, 4*S^2*D per attention layer."""
    return (2 * S * 1536 * 768 + 12 * (2 * S * 768 * 2304 + 4 * S * S * 768 + 2 * S * 768 * 768 + 4 * S * 768 * 3072)
            + (S // 4) * (2 * 3072 * 3072 + 2 * 3072 * 1024))


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"bf16_tflops": d.get("bf16_tflops_sustained", d.get("bf16_tflops")), "bf16_tflops_burst": d.get("bf16_tflops"),
                "hbm_gbs": d.get("hbm_gbs"), "source": "measured"}
    return {"bf16_tflops": 1400.0, "bf16_tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback"}


def measured_traffic():
    """DRAM bytes per launch of the dominant kernel family from the committed ncu --set full capture."""
    def order(f):   # r01_v10 after r01_v7: numeric, not lexicographic
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
# CPU path (oracle port of the reference) — baseline leg and `--impl reference`
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(px: int, sample_b: int, steps: int, warmup: int):
    from oracle import vision_oracle as VO

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = qwen_cfg(px)
    from llm_quest_b200.qwen.qwen3_5.qwen3_5_vision_model import Qwen3_5VisionModel

    torch.manual_seed(123)
    sd = {k: v.detach() for k, v in Qwen3_5VisionModel(cfg).state_dict().items()}  # parameter containers only
    pixels = torch.randn(sample_b, 3, 2, px, px, generator=torch.Generator().manual_seed(1234))
    times = []
    with torch.inference_mode():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            VO.qwen_vision_forward(sd, cfg, pixels)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    total = sum(times)
    return {"img_per_s": sample_b * len(times) / total, "ms_per_step": 1e3 * total / len(times), "cores": cores,
            "sample": f"{sample_b} images of {px}x{px} (T=2) per step, fp32, {len(times)} timed steps after {warmup} warm-up"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(args.px, args.cpu_batch, max(1, args.steps), max(1, min(args.warmup, 1)))
    line = {
        "impl": "reference", "metric": "images/sec (vision encode+fuse)", "value": round(r["img_per_s"], 3), "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(r["ms_per_step"], 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"cfg2: Qwen3.5 Qwen3-ViT tower + spatial-merge adapter, {args.px}x{args.px}, T=2 "
                               f"(CPU leg: bounded sample of {args.cpu_batch} images/step; no op mixes samples)"},
        "cpu_baseline": {"value": round(r["img_per_s"], 3), "unit": "images/s", "cores": r["cores"], "kind": "port",
                         "sample": r["sample"]},
        "e2e": {"value": round(r["img_per_s"], 3), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU path
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist

    from llm_quest_b200 import _lib, parallel
    from llm_quest_b200.qwen.qwen3_5.qwen3_5_vision_model import Qwen3_5VisionModel

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

    px, B = args.px, args.batch
    cfg = qwen_cfg(px)
    torch.manual_seed(123)
    model = Qwen3_5VisionModel(cfg).eval().to(dev)
    S = (px // 16) ** 2
    g = torch.Generator().manual_seed(1234 + rank)
    host_px = torch.randn(B, 3, 2, px, px, generator=g).to(torch.bfloat16).pin_memory()
    dev_px = host_px.to(dev, non_blocking=True)
    n_out = S // 4

    # N > 1: the all-gather is FUSED into the tower's last GEMM — its epilogue stores this rank's rows into every
    # rank's gathered buffer over NVLink peer memory (parallel.FusedAllGather), followed by a signal-pad barrier; no
    # NCCL kernel competes with the persistent kernels for SMs. If peer memory cannot be set up the NCCL all-gather
    # on a side stream is used instead (recorded in config.parallelism).
    fused, overlap, gather_kind = None, None, "none"
    if world > 1:
        try:
            fused = parallel.FusedAllGather(rows_local=B * n_out, cols=1024)
            gather_kind = ("all-gather fused into the last GEMM (" + ("NVSwitch multicast stores" if fused.multicast_ptr else "NVLink peer stores")
                           + " + signal barrier)")
        except Exception as e:  # noqa: BLE001
            print(f"[bench] fused all-gather unavailable ({type(e).__name__}: {e}); using NCCL", file=sys.stderr, flush=True)
            overlap = parallel.OverlappedAllGather(dev)
            gather_kind = "NCCL all-gather of merged embeddings on a side stream"

    def step_resident():
        if fused is not None:
            return model(dev_px, gather=fused)
        out = model(dev_px)
        if overlap is not None:
            out, _ = overlap.submit(out.to(torch.bfloat16))
        return out

    from llm_quest_b200.pipeline import StreamedEncoder

    # End to end at N > 1 the gathered batch stays in HBM (where the downstream LLM consumes it); the host read of the
    # step's result is each rank's own shard (every row of the job's output crosses PCIe once). Downloading the whole
    # gathered batch on every rank saturated host memory at 8 ranks: 18.6 k img/s against 38.6 k device-resident.
    gathered = [None]

    def gather_keep_local(o):
        local = o.to(torch.bfloat16)
        gathered[0] = overlap.submit(local)
        return local

    def fused_model(x):
        full = model(x, gather=fused)            # [world*B, n_out, 1024] bf16 on every rank
        gathered[0] = full
        return full[rank * B:(rank + 1) * B]

    if fused is not None:
        enc = StreamedEncoder(fused_model, depth=2, device=dev)
    else:
        enc = StreamedEncoder(model, depth=2, device=dev, post_fn=gather_keep_local if overlap is not None else None)
    sink = [0.0]

    def run_e2e(steps):
        """`steps` batches through the public streaming API: pinned-host pixels in, host embeddings out.
        Upload of batch i+1 and download of batch i-1 overlap the kernels of batch i."""
        for _ in range(steps):
            enc.submit(host_px)
            for out in enc.ready():
                sink[0] += float(out[0, 0, 0])   # touch the host result
        for out in enc.drain():
            sink[0] += float(out[0, 0, 0])

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

    with torch.inference_mode():
        for _ in range(max(3, args.warmup)):
            step_resident()
        _lib.reset_launch_count()
        with ClockSampler(local) as clocks:
            ms_total = timed(step_resident, args.steps)
        launches = _lib.launch_count()

        # per-kernel timing for the roofline (separate pass: event pairs around every launch)
        with _lib.KernelTimer() as kt:
            for _ in range(2):
                model(dev_px)
        torch.cuda.synchronize()
        fam = kt.summary()

        run_e2e(2)
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

    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total / 1e3)
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)
    peaks = measured_peaks()

    line = None
    if rank == 0:
        gemm_ms = sum(d["ms_total"] for k, d in fam.items() if k.startswith("gemm_") or k == "patch_embed")
        gemm_fl = sum(d["flops"] for k, d in fam.items() if k.startswith("gemm_") or k == "patch_embed")
        gemm_n = sum(d["launches"] for k, d in fam.items() if k.startswith("gemm_") or k == "patch_embed")
        all_ms = sum(d["ms_total"] for d in fam.values())
        achieved = gemm_fl / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
        breakdown = {}
        for k, d in sorted(fam.items(), key=lambda kv: -kv[1]["ms_total"]):
            e = {"launches": d["launches"] // 2, "ms_per_step": round(d["ms_total"] / 2, 3), "share": round(d["ms_total"] / all_ms, 3)}
            if d["flops"]:
                e["tflops"] = round(d["flops"] / (d["ms_total"] / 1e3) / 1e12, 1)
            if d["bytes"]:
                e["gbs"] = round(d["bytes"] / (d["ms_total"] / 1e3) / 1e9, 1)
            breakdown[k] = e
        step_tflops = tower_flops(S) * B / (ms_step / 1e3) / 1e12
        line = {
            "metric": "images/sec (vision encode+fuse)", "value": round(value, 1), "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": round(ms_step, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"cfg2: Qwen3.5 Qwen3-ViT tower + spatial-merge adapter, {px}x{px}, T=2, batch {B} per GPU",
                       "global_batch": world * B, "tokens_per_image": S, "l2_policy": "inputs_larger_than_l2 (154 MB pixels + 1.3 GB activations per step)",
                       "parallelism": f"sample-sharded x{world}" + (f" + {gather_kind}" if world > 1 else "")},
            "step_tflops": round(step_tflops, 1), "step_frac_of_peak": round(step_tflops / peaks["bf16_tflops"], 3),
            "roofline": {"bound": "tensor", "kernel": "vf::gemm_kernel<EPI,BN> (tcgen05, all epilogues; incl. patch-embed gather GEMM)",
                         "achieved": round(achieved, 1), "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                         "frac": round(achieved / peaks["bf16_tflops"], 3), "traffic": measured_traffic()[0],
                         "traffic_unit": "DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, avg over the family)",
                         "traffic_source": measured_traffic()[1], "peak_source": peaks["source"] + " (sustained cuBLAS bf16)",
                         "launches_per_step": gemm_n // 2, "avg_launch_ms": round(gemm_ms / max(gemm_n, 1), 4),
                         "share_of_step": round(gemm_ms / all_ms, 3)},
            "kernels": breakdown,
            "e2e": {"value": round(e2e_value, 1), "unit": "images/s", "ms_per_step": round(ms_e2e / args.steps, 3),
                    "h2d_bytes_per_step": world * host_px.numel() * host_px.element_size(),
                    "d2h_bytes_per_step": world * B * n_out * 1024 * (4 if world == 1 else 2),
                    "bytes_note": "whole job (all ranks); N>1: each rank reads back its own bf16 shard, the all-gathered batch stays in HBM",
                    "api": "llm_quest_b200.pipeline.StreamedEncoder(Qwen3_5VisionModel): pinned-host bf16 pixels in, merged "
                           "embeddings read back to pinned host memory every step; upload/compute/download on 3 streams"},
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
        }
        if world == 1 and not args.no_cpu:
            r = cpu_reference_run(px, args.cpu_batch, 2, 1)
            line["cpu_baseline"] = {"value": round(r["img_per_s"], 3), "unit": "images/s", "cores": r["cores"], "kind": "port",
                                    "sample": r["sample"]}
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
    ap.add_argument("--batch", type=int, default=64, help="images per GPU per step")
    ap.add_argument("--px", type=int, default=448)
    ap.add_argument("--cpu-batch", type=int, default=8, help="images per step of the CPU baseline sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

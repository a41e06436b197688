"""Host-side streaming runtime for the vision-encode path: overlap H2D, compute and D2H.

The reference's callers move a batch to the device, run the tower and read the result back in one
thread on one stream (qwen3_5_generate_multimodal.py:101-112); on a B200 the tower takes ~14 ms for
64 images while the PCIe copies take ~4 ms, so serialising them costs ~25 % of end-to-end
throughput. ``StreamedEncoder`` keeps the same per-batch semantics (host pixels in, host embeddings
out, in submission order) but runs three CUDA streams — upload, compute, download — over a small
ring of device/pinned-host buffers, with events carrying the dependencies. No thread, no graph
capture: just asynchronous copies and the kernels' own stream argument.

    enc = StreamedEncoder(model, depth=2)
    for batch in host_batches:           # pinned CPU tensors [B, C, T, H, W], bf16 or fp32
        enc.submit(batch)
        for out in enc.ready():          # finished results, in order (pinned CPU tensors)
            consume(out)
    for out in enc.drain():
        consume(out)
"""

from __future__ import annotations

from collections import deque

import torch


class GraphedEncoder:
    """Small-batch latency path: the tower's ~75 launches captured once in a CUDA graph and replayed.

    At batch 1-8 a forward is launch-bound (each kernel runs for a few microseconds); every libvfuse call takes raw
    pointers and a stream, allocates nothing and never synchronises, so the whole forward is capturable as is: the
    tensor maps are kernel parameters encoded at capture time against the graph's private buffers. Inputs are copied
    into the captured input buffer, the output buffer is returned (valid until the next call).

        g = GraphedEncoder(model, example_pixels)     # example fixes shape and dtype
        out = g(pixels)                               # same values as model(pixels), bit for bit
    """

    def __init__(self, model: torch.nn.Module, example: torch.Tensor, warmup: int = 2):
        if not example.is_cuda:
            raise RuntimeError("GraphedEncoder needs CUDA tensors (libvfuse has no CPU fallback)")
        self.model = model
        self.static_in = example.clone()
        side = torch.cuda.Stream(example.device)
        side.wait_stream(torch.cuda.current_stream(example.device))
        with torch.cuda.stream(side), torch.inference_mode():
            for _ in range(warmup):          # builds the packed-weight caches outside the capture
                model(self.static_in)
        torch.cuda.current_stream(example.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.inference_mode(), torch.cuda.graph(self.graph):
            self.static_out = model(self.static_in)

    def __call__(self, pixels: torch.Tensor) -> torch.Tensor:
        if pixels.shape != self.static_in.shape or pixels.dtype != self.static_in.dtype:
            raise ValueError(f"captured for {tuple(self.static_in.shape)} {self.static_in.dtype}, "
                             f"got {tuple(pixels.shape)} {pixels.dtype}")
        self.static_in.copy_(pixels, non_blocking=True)
        self.graph.replay()
        return self.static_out


class StreamedEncoder:
    def __init__(self, model: torch.nn.Module, depth: int = 2, device: torch.device | None = None, post_fn=None,
                 pre_fn=None):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.model = model
        self.post_fn = post_fn  # optional device-side step after the tower (e.g. an NCCL all-gather), same stream
        # optional device-side step before the tower, same stream: e.g. uint8 [B,H,W,3] uploads turned into pixel
        # tensors by qwen3_5.preprocess.pixels_from_uint8 (4x less PCIe traffic than bf16 (B,C,2,H,W) pixels)
        self.pre_fn = pre_fn
        self.depth = depth
        self.device = device if device is not None else next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("StreamedEncoder needs the model on a CUDA device (libvfuse has no CPU fallback)")
        self.s_up = torch.cuda.Stream(self.device)
        self.s_run = torch.cuda.Stream(self.device)
        self.s_down = torch.cuda.Stream(self.device)
        self._slots = [dict(dev_in=None, host_out=None, free=None) for _ in range(depth)]
        self._next = 0
        self._inflight = deque()  # (slot index, done event)

    def submit(self, host_pixels: torch.Tensor) -> None:
        """Enqueue one batch; blocks only when all `depth` slots are still in flight."""
        if len(self._inflight) == self.depth:
            raise RuntimeError("all slots busy: collect results with ready()/drain() before submitting more")
        i = self._next
        self._next = (self._next + 1) % self.depth
        slot = self._slots[i]
        if slot["dev_in"] is None or slot["dev_in"].shape != host_pixels.shape or slot["dev_in"].dtype != host_pixels.dtype:
            slot["dev_in"] = torch.empty(host_pixels.shape, dtype=host_pixels.dtype, device=self.device)
        ev_up, ev_run, ev_done = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
        with torch.cuda.stream(self.s_up):
            slot["dev_in"].copy_(host_pixels, non_blocking=True)
            ev_up.record()
        with torch.cuda.stream(self.s_run), torch.inference_mode():
            self.s_run.wait_event(ev_up)
            x = slot["dev_in"] if self.pre_fn is None else self.pre_fn(slot["dev_in"])
            out = self.model(x)
            if self.post_fn is not None:
                out = self.post_fn(out)
            ev_run.record()
        if slot["host_out"] is None or slot["host_out"].shape != out.shape or slot["host_out"].dtype != out.dtype:
            slot["host_out"] = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
        with torch.cuda.stream(self.s_down):
            self.s_down.wait_event(ev_run)
            slot["host_out"].copy_(out, non_blocking=True)
            out.record_stream(self.s_down)
            ev_done.record()
        self._inflight.append((i, ev_done))

    def ready(self, block_if_full: bool = True):
        """Yield finished results in submission order. When every slot is in flight (and
        block_if_full) wait for the oldest one, so the caller can always submit afterwards."""
        while self._inflight:
            i, ev = self._inflight[0]
            if not ev.query():
                if block_if_full and len(self._inflight) == self.depth:
                    ev.synchronize()
                else:
                    return
            self._inflight.popleft()
            yield self._slots[i]["host_out"]

    def drain(self):
        while self._inflight:
            i, ev = self._inflight.popleft()
            ev.synchronize()
            yield self._slots[i]["host_out"]

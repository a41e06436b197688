"""Host-side streaming runtime for the vision-encode path: overlap H2D, compute and D2H; replay steps from CUDA graphs.

The reference's callers move a batch to the device, run the tower and read the result back in one
thread on one stream (qwen3_5_generate_multimodal.py:101-112); on a B200 the tower takes ~12 ms for
64 images while the PCIe copies take ~6 ms, so serialising them costs a third of the end-to-end
throughput. ``StreamedEncoder`` keeps the same per-batch semantics (host inputs in, host results
out, in submission order) but runs three CUDA streams — upload, compute, download — over a small
ring of device/pinned-host buffers, with events carrying the dependencies. No thread: just asynchronous
copies, the kernels' own stream argument and (``graph=True``) one captured CUDA graph per ring slot.

    enc = StreamedEncoder(model, depth=2)
    for batch in host_batches:           # pinned CPU tensor(s): one tensor or a dict of tensors
        enc.submit(batch)
        for out in enc.ready():          # finished results, in order (tuples of pinned CPU tensors)
            consume(out)
    for out in enc.drain():
        consume(out)
"""

from __future__ import annotations

from collections import deque

import torch


class GraphedEncoder:
    """A pass of the path (~65-90 launches) captured once in a CUDA graph and replayed.

    Every libvfuse call takes raw pointers and a stream, allocates nothing and never synchronises, so a whole forward
    is capturable as is: the tensor maps are kernel parameters encoded at capture time against the graph's private
    buffers. At batch 1-8 a forward is launch-bound; at batch 64 the ~90 ctypes launches from Python still leave
    ~7 % of the step outside any kernel — the replay closes that gap.

        g = GraphedEncoder(model, example_pixels)     # example fixes shape and dtype
        out = g(pixels)                               # same values as model(pixels), bit for bit; valid until the next call
    or, for a step over fixed device buffers (bench.py):
        g = GraphedEncoder(lambda: step(), None); g.replay()
    """

    def __init__(self, model, example: torch.Tensor | None, warmup: int = 2, device=None):
        if example is not None and not example.is_cuda:
            raise RuntimeError("GraphedEncoder needs CUDA tensors (libvfuse has no CPU fallback)")
        self.static_in = example.clone() if example is not None else None
        fn = (lambda: model(self.static_in)) if example is not None else model
        dev = example.device if example is not None else (device or torch.device("cuda", torch.cuda.current_device()))
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.inference_mode():
            for _ in range(warmup):          # builds the packed-weight caches outside the capture
                fn()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.inference_mode(), torch.cuda.graph(self.graph):
            self.static_out = fn()

    def replay(self):
        self.graph.replay()
        return self.static_out

    def __call__(self, pixels: torch.Tensor) -> torch.Tensor:
        if self.static_in is None:
            raise ValueError("captured over fixed buffers: use replay()")
        if pixels.shape != self.static_in.shape or pixels.dtype != self.static_in.dtype:
            raise ValueError(f"captured for {tuple(self.static_in.shape)} {self.static_in.dtype}, "
                             f"got {tuple(pixels.shape)} {pixels.dtype}")
        self.static_in.copy_(pixels, non_blocking=True)
        self.graph.replay()
        return self.static_out


def bind_host_thread_to_gpu_numa(device_index: int) -> str:
    """Pin the calling process to the CPU cores NVML lists as local to GPU `device_index` (its NUMA node), so that the
    pinned host buffers allocated afterwards are first-touched on the memory next to that GPU's PCIe root. With eight
    ranks streaming 300 MB of pixels per step each, buffers that land on the other socket push every upload through
    the inter-socket link. Returns a short description of what was done (for the bench line); never raises."""
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return f"no narrower GPU-local CPU set ({len(allowed)} cores allowed)"
        os.sched_setaffinity(0, cpus)
        return f"bound to {len(cpus)} GPU-local cores of {len(allowed)}"
    except Exception as e:  # noqa: BLE001 — affinity is an optimisation, never a requirement
        return f"not bound ({type(e).__name__})"


def _as_dict(x):
    return x if isinstance(x, dict) else {"x": x}


def _as_tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x,)


class StreamedEncoder:
    """``model`` is called with what ``submit`` was given, moved to the device (a tensor, or a dict of tensors).

    graph=True: the compute of every ring slot is captured in its own CUDA graph the first time the slot sees a shape
    (device input buffers and outputs are then fixed per slot) and replayed afterwards.
    before_slot(i): optional hook called right before slot i's compute is enqueued or captured (e.g. to pin the
    destination slot of a fused all-gather to the ring slot)."""

    def __init__(self, model, depth: int = 2, device: torch.device | None = None, post_fn=None, pre_fn=None, graph: bool = False,
                 before_slot=None):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.model = model
        self.post_fn = post_fn  # optional device-side step after the tower (e.g. an NCCL all-gather), same stream
        # optional device-side step before the tower, same stream: e.g. uint8 [B,H,W,3] uploads turned into pixel
        # tensors by qwen3_5.preprocess.pixels_from_uint8 (4x less PCIe traffic than bf16 (B,C,2,H,W) pixels)
        self.pre_fn = pre_fn
        self.depth = depth
        self.graph = graph
        self.before_slot = before_slot
        self.device = device if device is not None else next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("StreamedEncoder needs the model on a CUDA device (libvfuse has no CPU fallback)")
        self.s_up = torch.cuda.Stream(self.device)
        self.s_run = torch.cuda.Stream(self.device)
        self.s_down = torch.cuda.Stream(self.device)
        self._slots = [dict(dev_in=None, host_out=None, sig=None, graph=None, out=None) for _ in range(depth)]
        self._next = 0
        self._inflight = deque()  # (slot index, done event)
        self.last_out_bytes = 0
        self.graph_failed = None

    def _compute(self, slot, single):
        x = slot["dev_in"]["x"] if single else slot["dev_in"]
        if self.pre_fn is not None:
            x = self.pre_fn(x)
        out = self.model(x)
        if self.post_fn is not None:
            out = self.post_fn(out)
        return _as_tuple(out)

    def submit(self, host_inputs) -> None:
        """Enqueue one batch; raises when all `depth` slots are still in flight."""
        if len(self._inflight) == self.depth:
            raise RuntimeError("all slots busy: collect results with ready()/drain() before submitting more")
        single = not isinstance(host_inputs, dict)
        hin = _as_dict(host_inputs)
        i = self._next
        self._next = (self._next + 1) % self.depth
        slot = self._slots[i]
        sig = tuple((k, tuple(v.shape), v.dtype) for k, v in hin.items())
        if slot["sig"] != sig:
            slot["dev_in"] = {k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in hin.items()}
            slot["sig"], slot["graph"], slot["out"] = sig, None, None
        ev_up, ev_run, ev_done = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
        with torch.cuda.stream(self.s_up):
            for k, v in hin.items():
                slot["dev_in"][k].copy_(v, non_blocking=True)
            ev_up.record()
        with torch.cuda.stream(self.s_run), torch.inference_mode():
            self.s_run.wait_event(ev_up)
            if self.before_slot is not None:
                self.before_slot(i)
            if self.graph and self.graph_failed is None and slot["graph"] is None:
                try:
                    g = torch.cuda.CUDAGraph()
                    self._compute(slot, single)             # packed-weight caches, allocator warm-up
                    self.s_run.synchronize()
                    if self.before_slot is not None:
                        self.before_slot(i)
                    with torch.cuda.graph(g, stream=self.s_run):
                        slot["out"] = self._compute(slot, single)
                    slot["graph"] = g
                except Exception as e:  # noqa: BLE001 — fall back to eager launches, keep the reason
                    self.graph_failed = f"{type(e).__name__}: {e}"
                    slot["graph"], slot["out"] = None, None
                    torch.cuda.synchronize()
            if slot["graph"] is not None:
                slot["graph"].replay()
                outs = slot["out"]
            else:
                outs = self._compute(slot, single)
            ev_run.record()
        if slot["host_out"] is None or any(h.shape != o.shape or h.dtype != o.dtype for h, o in zip(slot["host_out"], outs)):
            slot["host_out"] = tuple(torch.empty(o.shape, dtype=o.dtype, pin_memory=True) for o in outs)
        self.last_out_bytes = sum(o.numel() * o.element_size() for o in outs)
        with torch.cuda.stream(self.s_down):
            self.s_down.wait_event(ev_run)
            for h, o in zip(slot["host_out"], outs):
                h.copy_(o, non_blocking=True)
                o.record_stream(self.s_down)
            ev_done.record()
        self._inflight.append((i, ev_done, single))

    def _result(self, i, single):
        out = self._slots[i]["host_out"]
        return out[0] if single and len(out) == 1 else out

    def ready(self, block_if_full: bool = True):
        """Yield finished results in submission order. When every slot is in flight (and
        block_if_full) wait for the oldest one, so the caller can always submit afterwards."""
        while self._inflight:
            i, ev, single = self._inflight[0]
            if not ev.query():
                if block_if_full and len(self._inflight) == self.depth:
                    ev.synchronize()
                else:
                    return
            self._inflight.popleft()
            yield self._result(i, single)

    def drain(self):
        while self._inflight:
            i, ev, single = self._inflight.popleft()
            ev.synchronize()
            yield self._result(i, single)

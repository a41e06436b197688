"""Sample sharding across the GPUs of one box (one process per GPU, torch.distributed).

The vision-encode-and-fuse path has no cross-sample operation: rank r encodes
samples [r*B/G, (r+1)*B/G) with replicated weights and no communication. The only collective is the
all-gather (dim 0) of the fused embeddings — and of the [3, b, seq] position ids, gathered on
dim 1 — for a consumer that wants the whole batch on every rank. NCCL on GPUs, gloo in CPU tests.
"""

from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_samples: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced split: the first n % world ranks get one extra sample."""
    base, extra = divmod(n_samples, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(t: torch.Tensor, rank: int | None = None, world: int | None = None, dim: int = 0) -> torch.Tensor:
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_bounds(t.shape[dim], rank, world)
    return t.narrow(dim, lo, hi - lo)


def all_gather_cat(x: torch.Tensor, dim: int = 0, group=None) -> torch.Tensor:
    """Concatenate equally-shaped per-rank tensors along `dim` on every rank."""
    world = dist.get_world_size(group)
    if world == 1:
        return x
    xc = x.movedim(dim, 0).contiguous()
    out = torch.empty((world * xc.shape[0], *xc.shape[1:]), dtype=xc.dtype, device=xc.device)
    try:
        dist.all_gather_into_tensor(out, xc, group=group)
    except (RuntimeError, NotImplementedError):  # backends without the flat variant
        parts = list(out.chunk(world, dim=0))
        dist.all_gather(parts, xc, group=group)
    return out.movedim(0, dim)


def all_gather_fused(inputs_embs: torch.Tensor, position_ids: torch.Tensor | None = None, group=None):
    """inputs_embs [b_loc, seq, D] -> [b, seq, D]; position_ids [3, b_loc, seq] -> [3, b, seq]."""
    embs = all_gather_cat(inputs_embs, 0, group)
    pids = None if position_ids is None else all_gather_cat(position_ids, 1, group)
    return embs, pids


class OverlappedAllGather:
    """All-gather of step i on a side stream while step i+1 computes ("overlap per micro-batch on a
    side stream"). The collective is the only exchange of the sample-sharded path and its consumer (the downstream
    LLM) sits behind the whole encode step, so nothing in the next step has to wait for it.

        ag = OverlappedAllGather(device)
        for batch in batches:
            out = model(batch)                 # current stream
            gathered, done = ag.submit(out)    # enqueued behind `out`, runs beside the next model(...) call
        ag.wait_all()                          # current stream waits for every pending gather

    A consumer of `gathered` on another stream must wait for `done` (a CUDA event) first.
    """

    def __init__(self, device=None, depth: int = 2, dim: int = 0, group=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.stream = torch.cuda.Stream(self.device)
        self.depth, self.dim, self.group = depth, dim, group
        self._slots = [None] * depth      # keeps the last `depth` results (and their memory) alive
        self._n = 0

    def submit(self, x: torch.Tensor):
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            out = all_gather_cat(x, self.dim, self.group)
            x.record_stream(self.stream)
            done = torch.cuda.Event()
            done.record(self.stream)
        self._slots[self._n % self.depth] = (out, done)
        self._n += 1
        return out, done

    def wait_all(self):
        torch.cuda.current_stream(self.device).wait_stream(self.stream)


class FusedAllGather:
    """Destination of the all-gather that is fused into the producing GEMM (one B200 box, NVLink peer memory).

    Instead of "GEMM into a local buffer, then ncclAllGather", the last GEMM of the vision tower stores every output
    row straight into the gathered buffer of EVERY rank (``vf_epilogue.peer_out``: peer-mapped pointers from
    ``torch.distributed._symmetric_memory``), so the transfer rides on the epilogue tile by tile and no NCCL kernel
    competes with the persistent kernels for SMs. A signal-pad barrier after the launch makes the rows visible to
    every rank. Two slots alternate, so a consumer may still read step i while step i+1 is being written.

        fg = FusedAllGather(rows_local=B * n_merged, cols=1024)
        gathered = model(pixels, gather=fg)        # [world * rows_local, cols] on every rank
    """

    def __init__(self, rows_local: int, cols: int, dtype: torch.dtype = torch.bfloat16, group=None, slots: int = 2,
                 multicast: bool | None = None):
        import torch.distributed._symmetric_memory as symm

        self.group = dist.group.WORLD if group is None else group
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.rows_local, self.cols, self.dtype, self.slots = rows_local, cols, dtype, slots
        dev = torch.device("cuda", torch.cuda.current_device())
        self.buf = symm.empty((slots, self.world * rows_local, cols), dtype=dtype, device=dev)
        self.hdl = symm.rendezvous(self.buf, self.group)
        self._ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        # NVSwitch multicast (NVLS): ONE multimem.st to the multicast address is replicated by the switch into every
        # rank's buffer, so the producing SM sends each row once instead of `world` times (8 GPUs: 12.86 ms per step
        # against 13.06 with unicast peer stores and 13.19 with NCCL). Used whenever the fabric offers a multicast
        # address (multicast=False forces the unicast peer stores).
        mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
        want = True if multicast is None else bool(multicast)
        self.multicast_ptr = mc if (want and mc != 0) else 0
        self._slot = -1

    def set_next_slot(self, slot: int) -> None:
        """Pin the slot the next forward writes (a CUDA graph freezes it at capture: give every graph its own)."""
        self._slot = (slot - 1) % self.slots

    def next_slot(self) -> int:
        self._slot = (self._slot + 1) % self.slots
        return self._slot

    def peer_ptrs(self, slot: int) -> list[int]:
        """Address of THIS rank's first row inside slot `slot` of every rank's buffer (own buffer included)."""
        off = ((slot * self.world + self.rank) * self.rows_local * self.cols) * self.buf.element_size()
        if self.multicast_ptr:
            return [self.multicast_ptr + off]
        return [p + off for p in self._ptrs]

    def local_rows(self, slot: int) -> torch.Tensor:
        lo = self.rank * self.rows_local
        return self.buf[slot, lo:lo + self.rows_local]

    def gathered(self, slot: int) -> torch.Tensor:
        return self.buf[slot]

    def barrier(self) -> None:
        """Every rank's stores of the step have landed once every rank has passed this point (current stream)."""
        self.hdl.barrier()

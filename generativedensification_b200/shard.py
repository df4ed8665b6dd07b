"""Sharding of the per-object / per-view render loops over ranks (one process per GPU).

The reference renders every (object, view) pair in nested Python loops on one GPU
(lightning/network.py:813, 827-838, 848-856, 964-972; eval_all.py:3,9 pins
CUDA_VISIBLE_DEVICES=0).  Every render is independent given the object's Gaussians, so the
path shards with NO data-path collective:

  * batch sharding (primary):  objects round-robin over ranks; a rank holds only its objects.
  * view sharding (one huge object): Gaussians replicated, views strided `rank::world`.

NCCL (torch.distributed) is used only to gather scalars (loss, PSNR, timings) and, in the one
case where the views of a single object's densify vjp are split across ranks, to sum the
[P, 4] screen-space gradient (columns 2:4 are sums of absolute values, so they add too).
"""
from __future__ import annotations

import os
from typing import Dict, List, Sequence

import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world_size: int) -> List[int]:
    """Strided assignment: item i goes to rank i % world_size (mirrors a round-robin over the loop index)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    return list(range(rank, n_items, world_size))


def init_distributed(backend: str | None = None):
    """Initialise torch.distributed from the torchrun environment. Returns (rank, world_size, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local_rank


def world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def shutdown() -> None:
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


def barrier() -> None:
    if world() > 1:
        dist.barrier()


def gather_scalars(values: Sequence[float], device) -> torch.Tensor:
    """All-gather a short vector of fp32 scalars; returns [world, len(values)] on every rank."""
    t = torch.tensor(list(values), dtype=torch.float32, device=device).reshape(1, -1)
    if world() == 1:
        return t
    out = [torch.empty_like(t) for _ in range(world())]
    dist.all_gather(out, t)
    return torch.cat(out, dim=0)


def max_over_ranks(x: float, device) -> float:
    t = torch.tensor([x], dtype=torch.float64, device=device)
    if world() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(t: torch.Tensor) -> torch.Tensor:
    """In-place sum all-reduce (used for the [P, 4] screen-space gradient when views are split)."""
    if world() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def gather_view_results(local: Dict[int, torch.Tensor], n_views: int, device) -> torch.Tensor:
    """Collect per-view scalar results (e.g. PSNR) rendered under view sharding back in view order: a float32
    [n_views] tensor on every rank.  Ownership travels as an explicit count per view (a NaN result is a legitimate
    value, not a "not mine" marker); every view must have been rendered by exactly one rank."""
    vals = torch.zeros(n_views, dtype=torch.float32, device=device)
    owned = torch.zeros(n_views, dtype=torch.float32, device=device)
    for i, v in local.items():
        vals[i] = float(v)
        owned[i] = 1.0
    if world() > 1:
        both = torch.stack([vals, owned])
        gathered = [torch.empty_like(both) for _ in range(world())]
        dist.all_gather(gathered, both)
        stacked = torch.stack(gathered)                 # [world, 2, n_views]
        owned = stacked[:, 1].sum(0)
        mine = stacked[:, 1] > 0
        vals = torch.where(mine, stacked[:, 0], torch.zeros_like(stacked[:, 0])).sum(0)
    if int(owned.min()) != 1 or int(owned.max()) != 1:
        raise RuntimeError("view sharding did not cover every view exactly once")
    return vals


def broadcast_gaussians(gaussians: Dict[str, torch.Tensor], src: int = 0) -> Dict[str, torch.Tensor]:
    """View sharding of ONE object (BASELINE configs[4]; lightning/network.py:827-838 renders all views of an object
    from the same Gaussians): rank `src` holds the attributes, every rank ends up with a replica -- one NCCL broadcast
    per attribute tensor over NVLink (2 M Gaussians x 92 B = 184 MB), issued once per object.  Tensors must already be
    allocated with the right shapes on every rank (receivers may pass torch.empty)."""
    if world() > 1:
        for k in sorted(gaussians):
            dist.broadcast(gaussians[k], src=src)
    return gaussians


def gather_views(local: Dict[int, torch.Tensor], n_views: int, device=None) -> List[torch.Tensor]:
    """All-gather per-view tensors (e.g. rendered images) produced under view sharding (`rank::world`): returns the
    list of all n_views tensors in view order on every rank.  Every view must be present on exactly one rank."""
    if world() == 1:
        return [local[i] for i in range(n_views)]
    w, r = world(), dist.get_rank()
    some = next(iter(local.values())) if local else None
    if device is None:
        device = some.device if some is not None else torch.device("cuda", torch.cuda.current_device())
    shape = [torch.zeros(8, dtype=torch.int64, device=device)]
    if some is not None:
        shape[0][0] = some.dim()
        shape[0][1:1 + some.dim()] = torch.tensor(list(some.shape), device=some.device)
    shapes = [torch.empty_like(shape[0]) for _ in range(w)]
    dist.all_gather(shapes, shape[0])
    ref_shape = next(tuple(int(x) for x in s[1:1 + int(s[0])]) for s in shapes if int(s[0]) > 0)
    out: List[torch.Tensor] = [None] * n_views  # type: ignore[list-item]
    rounds = (n_views + w - 1) // w
    for j in range(rounds):  # round j: rank q contributes view j * w + q
        mine = j * w + r
        send = local[mine].contiguous() if mine < n_views else torch.zeros(ref_shape, dtype=torch.float32, device=device)
        recv = [torch.empty_like(send) for _ in range(w)]
        dist.all_gather(recv, send)
        for q in range(w):
            if j * w + q < n_views:
                out[j * w + q] = recv[q]
    return out


class HostFeeder:
    """Double-buffered host -> device feed of per-object Gaussian attribute dicts on a side stream.

    The reference feeds its render loops from a DataLoader and uploads each batch with `.to(device)` on
    the compute stream (lightning/network.py); here the upload of the next object overlaps the rendering
    of the current one.  submit(host_dict) enqueues an upload (pinned host tensors) into the next free
    slot; take() makes the compute stream wait for the oldest submitted slot and returns its device
    tensors; release() marks that slot reusable once the work enqueued so far on the compute stream is done."""

    def __init__(self, device, depth: int = 2):
        self.device = device
        self.depth = depth
        self.copy_stream = torch.cuda.Stream(device)
        self.slots = [None] * depth
        self.ready = [torch.cuda.Event() for _ in range(depth)]
        self.free = [None] * depth
        self.head = 0  # next slot to fill
        self.tail = 0  # next slot to take
        self.in_flight = 0
        self.taken = None

    def submit(self, host: Dict[str, torch.Tensor]) -> None:
        if self.in_flight >= self.depth:
            raise RuntimeError("HostFeeder: all slots are in flight; take()/release() first")
        i = self.head
        if self.slots[i] is None or any(self.slots[i][k].shape != v.shape for k, v in host.items()):
            self.slots[i] = {k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in host.items()}
            # The caching allocator may hand out blocks the compute stream freed while kernels that use them are still
            # queued there (reuse is stream-ordered on the ALLOCATING stream only): the copy stream must not write
            # them before that work has drained.  (Only on (re)allocation; afterwards the slot is ours.)
            self.copy_stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.copy_stream):
            if self.free[i] is not None:
                self.copy_stream.wait_event(self.free[i])  # the renders that read this slot have finished
            for k, v in host.items():
                self.slots[i][k].copy_(v, non_blocking=True)
            self.ready[i].record(self.copy_stream)
        self.head = (i + 1) % self.depth
        self.in_flight += 1

    def take(self) -> Dict[str, torch.Tensor]:
        if self.in_flight == 0:
            raise RuntimeError("HostFeeder: nothing submitted")
        i = self.tail
        torch.cuda.current_stream(self.device).wait_event(self.ready[i])
        self.taken = i
        self.tail = (i + 1) % self.depth
        return {k: v.detach() for k, v in self.slots[i].items()}

    def release(self) -> None:
        if self.taken is None:
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.free[self.taken] = ev
        self.taken = None
        self.in_flight -= 1

"""Data-parallel plumbing for the joint step: one process per GPU, volume pairs sharded across ranks, ONE
gradient all-reduce per step over a single flat contiguous fp32 bucket.

The reference has no distributed code at all; the unit of sharding is the (moving, fixed) pair
enumerated by ``RegDataSet*`` (lib/datasets.py:344-359).  BatchNorm statistics stay per replica (the reference
has no SyncBN and runs batch_size=1), so the gradient bucket is the only exchange step on the path.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed from torchrun's environment.  Returns (rank, local_rank, world_size)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


class FlatGradBucket:
    """All parameter gradients as views into ONE contiguous fp32 buffer.

    ``p.grad`` of every parameter is pre-set to a view of ``self.flat`` so autograd accumulates straight
    into the bucket; ``zero()`` is one memset, ``allreduce()`` one collective (sum, then 1/world scaling)."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatGradBucket: no trainable parameters")
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def zero(self):
        self.flat.zero_()

    def allreduce(self, world: int | None = None):
        """One collective per step.  No-op for a single process."""
        if not dist.is_available() or not dist.is_initialized():
            return
        world = world or dist.get_world_size()
        if world == 1:
            return
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        self.flat.mul_(1.0 / world)


def broadcast_parameters(module: torch.nn.Module, src: int = 0):
    """Make every replica start from rank ``src``'s weights and buffers (called once, outside the step)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src)


def shard_pairs(num_items: int, rank: int, world: int, seed: int = 230, epoch: int = 0):
    """Rank-strided slice of the shuffled ordered-pair index list.  Pair id -> (fixed, moving) follows
    lib/datasets.py:344-359: fixed = id // (N-1); moving = id % (N-1), +1 if >= fixed."""
    n_pairs = num_items * (num_items - 1)
    g = torch.Generator().manual_seed(seed + epoch)
    perm = torch.randperm(n_pairs, generator=g).tolist()
    out = []
    for pid in perm[rank::world]:
        fixed = pid // (num_items - 1)
        moving = pid % (num_items - 1)
        if moving >= fixed:
            moving += 1
        out.append((moving, fixed))
    return out

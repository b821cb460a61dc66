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
    into the bucket; ``zero()`` is one memset, ``allreduce()`` one collective (sum, then 1/world scaling).

    The reference trainer calls ``optimizer.zero_grad()`` (models/segmentation.py:142), whose default
    ``set_to_none=True`` drops the views: the next backward would then allocate fresh gradients OUTSIDE the bucket and
    ``allreduce()`` would average zeros.  ``rebind()`` -- called by ``allreduce()`` -- detects that (``p.grad`` is None
    or does not alias its slot), copies such gradients into the bucket and restores the views, so the collective is
    always over the real gradients.  Prefer ``zero()`` / ``zero_grad()`` of the bucket (one memset, views kept)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], inplace: bool = True):
        """``inplace``: the CUDA backward kernels add parameter gradients straight into the bucket views (ops._grad_dst)
        instead of returning tensors for autograd to add -- one launch less per parameter and use.  Turn it off for code
        that calls ``torch.autograd.grad`` on these parameters or hangs hooks on them."""
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatGradBucket: no trainable parameters")
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self._views: List[torch.Tensor] = []
        off = 0
        for p in self.params:
            n = p.numel()
            v = self.flat[off:off + n].view_as(p)
            if inplace and dev.type == "cuda":
                v._da_inplace = True
            self._views.append(v)
            p.grad = v
            off += n
        self.rebound = 0   # gradients found outside the bucket so far (diagnostic)
        self.flat_alt = None

    def enable_alt(self):
        """A second flat buffer with the same layout, reachable as ``p.grad._da_alt``: ops created under
        ``ops.grad_slot(1)`` add their parameter gradients there (a second pass of the same network on another stream);
        ``fold_alt()`` -- called by ``allreduce()`` -- adds it into the primary buffer."""
        if self.flat_alt is None:
            self.flat_alt = torch.zeros_like(self.flat)
            off = 0
            for p, v in zip(self.params, self._views):
                n = p.numel()
                v._da_alt = self.flat_alt[off:off + n].view_as(p)
                off += n
        return self

    def fold_alt(self):
        if self.flat_alt is not None:
            self.flat.add_(self.flat_alt)
            self.flat_alt.zero_()

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def zero(self):
        self.flat.zero_()
        if self.flat_alt is not None:
            self.flat_alt.zero_()
        self.rebind(copy=False)

    zero_grad = zero

    def rebind(self, copy: bool = True) -> int:
        """Make every ``p.grad`` a view of its bucket slot again.  With ``copy`` a gradient that autograd produced
        outside the bucket (after ``optimizer.zero_grad(set_to_none=True)``) is copied into its slot first; a
        parameter without a gradient gets a zeroed slot.  Returns the number of parameters that had to be fixed."""
        fixed = 0
        for p, v in zip(self.params, self._views):
            g = p.grad
            if g is not None and g.data_ptr() == v.data_ptr() and g.shape == v.shape and g.is_contiguous():
                continue
            fixed += 1
            if copy:
                if g is None:
                    v.zero_()
                else:
                    v.copy_(g)
            p.grad = v
        self.rebound += fixed
        return fixed

    def allreduce(self, world: int | None = None):
        """One collective per step.  No collective for a single process (the views are still re-checked).  Weight
        gradients that ``ops.set_wgrad_overlap`` put on the side stream are joined first."""
        if self.flat.is_cuda:
            from . import ops
            ops.join_wgrad_stream()
        self.fold_alt()
        if not dist.is_available() or not dist.is_initialized():
            self.rebind(copy=True)
            return
        world = world or dist.get_world_size()
        self.rebind(copy=True)   # every rank does this before the collective: the bucket always holds the real gradients
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


def shard_pairs(num_items: int, rank: int, world: int, seed: int = 230, epoch: int = 0, drop_last: bool = False):
    """Rank-strided slice of the shuffled ordered-pair index list.  Pair id -> (fixed, moving) follows
    lib/datasets.py:344-359: fixed = id // (N-1); moving = id % (N-1), +1 if >= fixed.

    EVERY rank gets the same number of pairs (one gradient all-reduce per pair: unequal counts would leave some
    ranks waiting in NCCL for a collective the others never issue).  When N(N-1) is not a multiple of ``world`` the
    shuffled list is padded by wrapping around, as torch's DistributedSampler does (``drop_last=True`` truncates
    instead); ``len(shard_pairs(...)) == ceil(N(N-1) / world)`` (floor with ``drop_last``) on all ranks."""
    if not (0 <= rank < world):
        raise ValueError(f"shard_pairs: rank {rank} outside world of {world}")
    n_pairs = num_items * (num_items - 1)
    g = torch.Generator().manual_seed(seed + epoch)
    perm = torch.randperm(n_pairs, generator=g).tolist()
    if drop_last:
        perm = perm[:(n_pairs // world) * world]
    elif n_pairs % world and n_pairs:
        pad = world - n_pairs % world
        perm = perm + (perm * ((pad + n_pairs - 1) // n_pairs))[:pad]
    out = []
    for pid in perm[rank::world]:
        fixed = pid // (num_items - 1)
        moving = pid % (num_items - 1)
        if moving >= fixed:
            moving += 1
        out.append((moving, fixed))
    return out

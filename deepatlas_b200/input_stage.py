"""Device-side input stage.

The reference's trainer moves every sample with ``images.cuda()`` and ``truths.long().cuda()`` on the compute
stream, from pageable memory, after clipping to [0, 1] (``SitkToTensor``, lib/transforms.py:79-80) and cropping
(``CropTensor``, lib/transforms.py:124-158) on the CPU (models/segmentation.py:152-154).  Here the RAW volume goes
through a pinned staging buffer and an asynchronous copy on a side stream, one step ahead of the compute stream
(double buffering), and the crop + clip run on the device (csrc/variants.cu).  Labels stay uint8 end to end: every
loss of this package reads uint8 labels directly, so ``.long()`` (8x the bytes) and the one-hot never exist.
"""
from __future__ import annotations

from typing import Iterable, Iterator, Optional, Sequence, Tuple

import torch

from . import ops


def crop_window(shape: Sequence[int], crop_size: Optional[Sequence[int]]) -> Tuple[Tuple[int, int, int], Tuple[int, int, int]]:
    """``CropTensor`` arithmetic (lib/transforms.py:129-157): a length-3 ``crop_size`` is removed from both sides of
    each axis, a length-6 one is (low_d, low_h, low_w, high_d, high_h, high_w).  Returns (low corner, size)."""
    D, H, W = (int(v) for v in shape[-3:])
    if crop_size is None:
        return (0, 0, 0), (D, H, W)
    crop_size = list(crop_size)
    if len(crop_size) == 3:
        crop_size = crop_size + crop_size
    elif len(crop_size) != 6:
        raise ValueError("crop size should be of length 3 or 6, but {} is given".format(len(crop_size)))
    lo = (crop_size[0], crop_size[1], crop_size[2])
    size = (D - crop_size[3] - lo[0], H - crop_size[4] - lo[1], W - crop_size[5] - lo[2])
    if min(size) < 1:
        raise ValueError(f"crop {crop_size} leaves nothing of a {(D, H, W)} volume")
    return lo, size


class DeviceInputStage:
    """Pinned double-buffered host-to-device copies plus crop / clip on the device.

    ``submit(image, seg)`` takes CPU tensors (image float32 ``(..., D, H, W)``, seg uint8 ``(..., D, H, W)`` or None),
    starts their copies on the side stream and returns immediately; ``get()`` returns the oldest submitted sample as
    device tensors, ordered after the copies on the caller's current stream."""

    def __init__(self, device, crop_size=None, clip=(0.0, 1.0), depth: int = 2):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("deepatlas_b200: the input stage needs a CUDA device (no CPU path exists)")
        self.crop_size, self.clip, self.depth = crop_size, clip, int(depth)
        self.stream = torch.cuda.Stream(device=self.device)
        self._slots = [dict() for _ in range(self.depth)]
        self._queue = []
        self._next = 0
        self.h2d_bytes = 0

    def _pinned(self, slot, key, like):
        buf = slot.get(key)
        if buf is None or buf.shape != like.shape or buf.dtype != like.dtype:
            buf = torch.empty(like.shape, dtype=like.dtype, pin_memory=True)
            slot[key] = buf
        return buf

    def submit(self, image: torch.Tensor, seg: Optional[torch.Tensor] = None):
        if len(self._queue) >= self.depth:
            raise RuntimeError("DeviceInputStage: all staging slots are in flight; call get() first")
        if image.dtype != torch.float32:
            image = image.float()
        if seg is not None and seg.dtype != torch.uint8:
            seg = seg.to(torch.uint8)
        slot = self._slots[self._next]
        self._next = (self._next + 1) % self.depth
        ev = slot.get("free")
        if ev is not None:
            ev.synchronize()  # the copy that last read this slot's pinned buffers has finished
        # a tensor that already sits in pinned memory is copied from where it is (no second host copy); the caller
        # must then leave it alone until the copy has run -- get() of the same sample orders the consumer after it
        if image.is_pinned() and image.is_contiguous():
            pi = image
        else:
            pi = self._pinned(slot, "image", image)
            pi.copy_(image)
        ps = None
        if seg is not None:
            if seg.is_pinned() and seg.is_contiguous():
                ps = seg
            else:
                ps = self._pinned(slot, "seg", seg)
                ps.copy_(seg)
        lo, size = crop_window(image.shape, self.crop_size)
        with torch.cuda.stream(self.stream):
            di = pi.to(self.device, non_blocking=True)
            ds = ps.to(self.device, non_blocking=True) if ps is not None else None
            free = torch.cuda.Event()
            free.record(self.stream)
            slot["free"] = free
            di = ops.crop_clip(di, lo, size, self.clip)
            if ds is not None and size != tuple(ds.shape[-3:]):
                ds = ops.crop_labels(ds, lo, size)
            ready = torch.cuda.Event()
            ready.record(self.stream)
        self.h2d_bytes += pi.numel() * 4 + (ps.numel() if ps is not None else 0)
        self._queue.append((di, ds, ready))

    def get(self):
        if not self._queue:
            raise RuntimeError("DeviceInputStage: nothing submitted")
        di, ds, ready = self._queue.pop(0)
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ready)
        di.record_stream(cur)
        if ds is not None:
            ds.record_stream(cur)
        return di, ds


def prefetch(loader: Iterable, device, crop_size=None, clip=(0.0, 1.0)) -> Iterator:
    """Wrap a loader that yields ``(images, truths, name)`` CPU batches (lib/datasets.py:68 through the DataLoader at
    models/segmentation.py:71-76): yields ``(images_dev, truths_dev_uint8, name)`` with the next batch's copies
    already in flight while the current one is being consumed."""
    stage = DeviceInputStage(device, crop_size=crop_size, clip=clip, depth=2)
    names = []
    for images, truths, name in loader:
        stage.submit(images, truths)
        names.append(name)
        if len(names) == 2:
            di, ds = stage.get()
            yield di, ds, names.pop(0)
    while names:
        di, ds = stage.get()
        yield di, ds, names.pop(0)

"""GPU evaluation path: the step the reference runs right after the hot path every epoch.

Mirror of the inner loop of ``SegmentationExperiment.eval`` (models/segmentation.py:185-194): for every volume,
``torch.max(pred, 1)[1]`` and, for each class c >= 1, ``metricEval('dice', argmax == c, truth == c, num_labels=2)``
(= ``1 - scipy.spatial.distance.dice`` on the two boolean arrays, lib/evalMetrics.py:58-68).  One CUDA pass produces the
label map and exact integer overlap counts; the closing division is done in float64 like scipy's.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from .ops import _KIND, _f32, _p, _stream


def argmax_counts(logits: torch.Tensor, truths: torch.Tensor | None = None, want_pred: bool = True):
    """Returns (counts int64 [N,3,C] = (#argmax==c, #truth==c, #both), pred uint8 [N,D,H,W] or None)."""
    logits = _f32(logits, "logits")
    N, C = logits.shape[:2]
    V = logits[0, 0].numel()
    kind = 0
    if truths is not None:
        if not truths.is_cuda or truths.is_floating_point():
            raise RuntimeError("deepatlas_b200: 'truths' must be an integer CUDA tensor")
        if truths.dtype not in _KIND:
            truths = truths.long()
        truths = truths.contiguous()
        if truths.numel() != N * V:
            raise ValueError("argmax_counts: truths must have N*D*H*W elements")
        kind = _KIND[truths.dtype]
    counts = torch.empty((N, 3, C), dtype=torch.int64, device=logits.device)
    pred = torch.empty((N,) + tuple(logits.shape[2:]), dtype=torch.uint8, device=logits.device) if want_pred else None
    _lib.call("da_argmax_counts", _p(logits), _p(truths), kind, N, C, V, _p(counts), _p(pred), _stream())
    return counts, pred


def dice_per_class(logits: torch.Tensor, truths: torch.Tensor) -> torch.Tensor:
    """(N, C-1) float64: 1 - scipy.spatial.distance.dice(argmax == c, truth == c) for c = 1..C-1.  As in scipy, a class
    absent from both the prediction and the truth gives nan (0/0)."""
    counts, _ = argmax_counts(logits, truths, want_pred=False)
    c = counts.double()
    P, T, I = c[:, 0, 1:], c[:, 1, 1:], c[:, 2, 1:]
    # scipy: dice = (n_tf + n_ft) / (2 n_tt + n_tf + n_ft) in float64, then 1 - dice; same operation order here
    diff = P + T - 2.0 * I
    return 1.0 - diff / (2.0 * I + diff)


_LKIND = dict(_KIND)
_LKIND[torch.float32] = 4


def label_overlap_counts(a: torch.Tensor, b: torch.Tensor, bins: int = 256) -> torch.Tensor:
    """counts int64 [N,3,bins] = (#[a == c], #[b == c], #[a == b == c]) for two label maps with the same number of
    elements per sample (integer dtypes, or float32 label values, truncated like ``mask.long()``)."""
    if not a.is_cuda or not b.is_cuda:
        raise RuntimeError("deepatlas_b200: label maps must be CUDA tensors (no CPU fallback exists)")
    if a.shape != b.shape:
        raise AssertionError("label maps must have the same shape")

    def prep(t):
        if t.dtype not in _LKIND:
            t = t.float() if t.is_floating_point() else t.long()
        return t.contiguous()
    a, b = prep(a), prep(b)
    N = a.shape[0]
    V = a[0].numel()
    counts = torch.empty((N, 3, bins), dtype=torch.int64, device=a.device)
    _lib.call("da_label_overlap_counts", _p(a), _LKIND[a.dtype], _p(b), _LKIND[b.dtype], N, int(bins), V, _p(counts), _stream())
    return counts

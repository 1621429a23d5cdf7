"""Host side of the converters (reference: convert.swift:59-127 `convertMistral`, q4_convert.py:41-81).

FP16: `ops.bucketize` is one C-ABI call.  Q4: the reference first pulls the global top-2 % of |w| out of the
matrix as fp32 outlier records (q4_draft.py:71-105) -- a one-off global sort, done here with torch on the device --
then bucketizes the remainder (effort_q4_bucketize)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib, ops
from ._lib import check


def q4_convert(core2: torch.Tensor, perc: float = 0.02) -> dict:
    """core2: W^T [in, out] float16 CUDA (q4_convert.py:53 passes `.T`).  Returns the tensors of
    q4_draft.convert(): probes, bucket.stats [in*8,2] f32, buckets [in*8,out/32] (f16 view of the words),
    outliers [N,4] f32 (value, in, out, 0)."""
    L = _lib.load()
    ops._need(core2, torch.float16, "core2")
    inn, out = core2.shape
    core = core2.clone()
    flat = core.view(-1)
    top = int(flat.numel() * perc)
    # argsort(-|w|)[:top] (q4_draft.py:79); ties at the 2 % boundary are UNPINNED (numpy quicksort is unstable).
    # One-time convert step, off the hot path: the outlier pick uses the library sort (torch.sort), everything after it
    # (bucketing, stats, probes) is this repo's kernels (csrc/q4.cuh).
    order = torch.sort(flat.abs().float(), descending=True, stable=True).indices[:top]
    outliers = torch.zeros((top, 4), dtype=torch.float32, device=core.device)
    outliers[:, 0] = flat[order].float()
    outliers[:, 1] = (order // out).float()
    outliers[:, 2] = (order % out).float()
    flat[order] = 0
    t = q4_bucketize(core)
    t["outliers"] = outliers
    return t


def q4_bucketize(core: torch.Tensor) -> dict:
    """The bucketize part of q4_draft.convert (:107-322) on W^T [in,out] with the outliers already zeroed."""
    L = _lib.load()
    ops._need(core, torch.float16, "core")
    inn, out = core.shape
    buckets = torch.empty((inn * 8, out // 32), dtype=torch.float16, device=core.device)
    stats = torch.empty((inn * 8, 2), dtype=torch.float32, device=core.device)
    probes = torch.empty((min(inn, out),), dtype=torch.float16, device=core.device)
    check(L.effort_q4_bucketize(core.data_ptr(), inn, out, buckets.data_ptr(), stats.data_ptr(), probes.data_ptr(),
                                ops._stream_ptr()), "effort_q4_bucketize")
    return {"probes": probes, "bucket.stats": stats, "buckets": buckets}

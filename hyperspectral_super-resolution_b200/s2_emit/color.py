"""Shared percentile stretch — call surface of the reference's ``s2_emit/color.py``
(``apply_shared_percentile_stretch`` :25-34, the step between the SRF synthesis and the polynomial fit at
``s2_emit/poly_regression.py:126-127``).

The percentiles are exact (three-pass radix select on the GPU, numpy's "linear" interpolation reproduced
operation by operation), so the stretched image is bit-identical to the reference's.  Histogram matching
and the Sinkhorn/OT colour transfer of the same reference file (:36-116, third-party POT) are outside the
hot path and not provided.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import kernels
from .._host import is_numpy_like, to_device, to_host


def shared_percentile_limits(img, mask, pmin: float = 2, pmax: float = 98):
    """(C, 2) float64: per channel ``np.percentile(img[..., c][mask], [pmin, pmax])`` (color.py:30-32).
    img: (H, W, C); mask: (H, W) bool.  Raises IndexError when the mask selects nothing, as numpy does."""
    numpy_in = is_numpy_like(img)
    x = to_device(img, torch.float32)
    if x.dim() != 3:
        raise ValueError(f"img must be (H, W, C), got {tuple(x.shape)}")
    m = to_device(mask, torch.uint8, x.device)
    if m.shape != x.shape[:2]:
        raise IndexError(f"boolean index did not match: mask {tuple(m.shape)} vs image {tuple(x.shape[:2])}")
    planes = kernels.alloc_planes(x.shape[2], x.shape[:2], x.device)
    planes.copy_(x.permute(2, 0, 1))
    lohi = kernels.masked_percentiles(planes, m, [pmin, pmax]).view(x.shape[2], 2)
    if numpy_in:
        if not bool(m.any()):
            raise IndexError("index -1 is out of bounds for axis 0 with size 0")   # np.percentile of nothing
        return to_host(lohi), planes
    return lohi, planes


def apply_shared_percentile_stretch(img, mask, pmin: float = 2, pmax: float = 98):
    """Per-channel percentile stretch within ``mask``; float32 in [0, 1] (reference :25-34):
    ``out[..., c] = clip((img[..., c] - lo) / (hi - lo + 1e-12), 0, 1)`` with
    ``lo, hi = np.percentile(img[..., c][mask], [pmin, pmax])``.  numpy in -> numpy out, CUDA in -> CUDA out."""
    numpy_in = is_numpy_like(img)
    lohi, planes = shared_percentile_limits(img, mask, pmin, pmax)
    if numpy_in:
        lohi = to_device(lohi, torch.float64, planes.device)
    C = planes.shape[0]
    out = kernels.stretch_apply(planes, lohi.view(C, 1, 2))
    out = out.permute(1, 2, 0).contiguous()
    return to_host(out, np.float32) if numpy_in else out

"""Shared percentile stretch — call surface of the reference's ``s2_emit/color.py``
(``apply_shared_percentile_stretch`` :25-34, the step between the SRF synthesis and the polynomial fit at
``s2_emit/poly_regression.py:126-127``).

The percentiles are exact (three-pass radix select on the GPU, numpy's "linear" interpolation reproduced
operation by operation), so the stretched image is bit-identical to the reference's.
``ot_match_rgb_sinkhorn_pot`` (:63-116) — the 3-D colour transfer: Sinkhorn OT on masked samples, barycentric targets,
affine fit, apply — reuses the OT kernels of ``fit_ot_poly_rgb`` plus an affine fit / apply pair (POT itself is neither
vendored nor pinned by the reference: the kernels follow its published ``dist`` / ``sinkhorn_knopp``).
``histogram_match_rgb`` (:36-63): CDF matching on sorted masked samples (library sort + our look-up kernels).  ``robust_norm`` / ``robust_norm_rgb``
(:6-23) are the float64 variants of the stretch.
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


def ot_match_rgb_sinkhorn_pot(src_rgb, ref_rgb, mask, n_samples: int = 5_000, reg: float = 0.05, numItermax: int = 300,
                              stopThr: float = 1e-6, seed: int = 0):
    """3-D colour transfer using Sinkhorn OT on masked RGB samples + affine fit (reference :63-116, same signature):
    masked rows with every channel finite (src and ref filtered independently, :83-87), fewer than 2 of either -> a copy
    of ``src_rgb`` (:89-90), ``default_rng(seed).choice`` samples (drawn on the host with numpy's generator, gathered
    on the GPU), squared-euclidean cost + Sinkhorn + barycentric targets (:98-104), ``lstsq([X 1], Ybar)`` (:106-109)
    and ``out[mask] = clip(out[mask] @ A + t, 0, 1)`` on the float32 copy (:111-116) — all on the GPU in fp64."""
    numpy_in = is_numpy_like(src_rgb)
    src = to_device(src_rgb, torch.float32)
    ref = to_device(ref_rgb, torch.float32, src.device)
    if src.shape != ref.shape or src.dim() != 3:
        raise ValueError(f"src_rgb / ref_rgb must be (H,W,C) of equal shape, got {tuple(src.shape)}, {tuple(ref.shape)}")
    m = to_device(mask, torch.uint8, src.device)
    if tuple(m.shape) != tuple(src.shape[:2]):
        raise IndexError(f"boolean index did not match: mask {tuple(m.shape)} vs image {tuple(src.shape[:2])}")
    C = src.shape[2]
    x2, y2 = src.reshape(-1, C), ref.reshape(-1, C)
    idx_x, nx = kernels.compact_finite_rows(x2, m.reshape(-1))
    idx_y, ny = kernels.compact_finite_rows(y2, m.reshape(-1))
    nx, ny = int(nx.item()), int(ny.item())                        # the sample draw needs them on the host
    if nx < 2 or ny < 2:                                           # :89-90
        return np.array(src_rgb, copy=True) if numpy_in else src_rgb.clone()
    rng = np.random.default_rng(seed)                              # :78
    ns, nt = min(int(n_samples), nx), min(int(n_samples), ny)
    sel_x = torch.from_numpy(rng.choice(nx, size=ns, replace=False).astype(np.int64)).to(src.device)   # :95
    sel_y = torch.from_numpy(rng.choice(ny, size=nt, replace=False).astype(np.int64)).to(src.device)   # :96
    X = kernels.gather_rows_f64(x2, idx_x, sel_x)
    Y = kernels.gather_rows_f64(y2, idx_y, sel_y)
    ybar, _ = kernels.sinkhorn_barycentric(X, Y, reg, numItermax, stopThr)                              # :98-104
    W = kernels.affine_fit(X, ybar)                                                                     # :106-109
    out = kernels.affine_apply(src, W, m, lo=0.0, hi=1.0)                                               # :111-116
    return to_host(out, np.float32) if numpy_in else out


def robust_norm(x, pmin: float = 2, pmax: float = 98):
    """``clip((x - lo) / (hi - lo + 1e-12), 0, 1)`` with ``lo, hi = np.nanpercentile(x, [pmin, pmax])`` over the whole
    array (reference :6-8): float64 of x's shape, NaN stays NaN.  Exact percentiles (radix select over the non-NaN
    samples), float64 stretch — bit-identical to the reference for float32 input."""
    numpy_in = is_numpy_like(x)
    t = to_device(x, torch.float32).contiguous()
    flat = t.reshape(1, -1)
    keep = kernels.notnan_mask(flat)
    lohi = kernels.masked_percentiles(flat, keep, [pmin, pmax])                       # [1, 1, 2]
    out = kernels.stretch_apply_f64(flat, lohi.view(1, 1, 2)).view(t.shape)
    return to_host(out, np.float64) if numpy_in else out


def robust_norm_rgb(img, mask, pmin: float = 2, pmax: float = 98):
    """Per-channel percentile stretch within ``mask``; float64, NaN outside the mask (reference :10-23)."""
    numpy_in = is_numpy_like(img)
    lohi, planes = shared_percentile_limits(img, mask, pmin, pmax)
    if numpy_in:
        lohi = to_device(lohi, torch.float64, planes.device)
    C = planes.shape[0]
    m = to_device(mask, torch.uint8, planes.device)
    out = kernels.stretch_apply_f64(planes, lohi.view(C, 1, 2), m.reshape(-1))
    out = out.permute(1, 2, 0).contiguous()
    return to_host(out, np.float64) if numpy_in else out


def histogram_match_rgb(src_rgb, ref_rgb, mask):
    """Histogram-match each channel independently within ``mask``; inputs assumed in [0, 1] and finite (reference
    :55-63, ``_hist_match_channel`` :36-53): masked source samples go through the source CDF and the inverse reference
    CDF built from ``np.unique`` counts, everything is clipped to [0, 1].  float32; numpy in -> numpy out."""
    numpy_in = is_numpy_like(src_rgb)
    src = to_device(src_rgb, torch.float32)
    ref = to_device(ref_rgb, torch.float32, src.device)
    if src.shape != ref.shape or src.dim() != 3:
        raise ValueError(f"src_rgb / ref_rgb must be (H,W,C) of equal shape, got {tuple(src.shape)}, {tuple(ref.shape)}")
    m = to_device(mask, torch.uint8, src.device)
    if tuple(m.shape) != tuple(src.shape[:2]):
        raise IndexError(f"boolean index did not match: mask {tuple(m.shape)} vs image {tuple(src.shape[:2])}")
    out = torch.empty_like(src)
    for c in range(src.shape[2]):
        out[..., c] = kernels.hist_match_channel(src[..., c], ref[..., c], m)
    return to_host(out, np.float32) if numpy_in else out

"""Per-channel polynomial colour matching — call surface of the reference's
``s2_emit/poly_regression.py`` (``fit_ot_poly_rgb`` :16-62, ``apply_poly_rgb`` :65-84).

The import-time script of the reference (:86-172, hard-coded /content paths) is deliberately not
reproduced.  The least-squares core (np.polyfit at :58-60) runs as fp64 normal-equation moments
+ a warp-level solve on the GPU; the apply is a fused Horner + mask + clip kernel.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import kernels
from .._host import is_numpy_like, to_device, to_host


def poly_fit(x, y, mask, deg: int = 2, *, min_count: int = 0):
    """Pixel-paired per-band fit: ``coeffs[k] = np.polyfit(x[k][m], y[k][m], deg)`` with
    ``m = mask & isfinite(x[k]) & isfinite(y[k])`` over ALL pixels (the per-band calibration of
    Pairs_EMIT_S2_demo-2.ipynb cell 72 generalised to degree ``deg``).

    x, y: (K, H, W) planes; mask: (H, W) or (K, H, W) bool or None.  Returns (K, deg+1) float64,
    highest power first — numpy in -> numpy out, CUDA tensors in -> CUDA tensor out.
    Bands with fewer than ``min_count`` samples get the identity polynomial.
    """
    numpy_in = is_numpy_like(x)
    xt = to_device(x, torch.float32)
    yt = to_device(y, torch.float32, xt.device)
    mt = None if mask is None else to_device(mask, torch.uint8, xt.device)
    coeffs = kernels.poly_fit(xt, yt, mt, deg, min_count=min_count)
    return to_host(coeffs) if numpy_in else coeffs


def fit_ot_poly_rgb(src_rgb, ref_rgb, mask, deg=2, n_samples=5000, reg=0.05, numItermax=300, stopThr=1e-6,
                    seed=0, *, targets="ot"):
    """Fit per-channel polynomial mapping y = poly(x); returns (3, deg+1) float64, highest power first.

    Signature of the reference (:16-24) plus ``targets``:
      * ``"paired"`` — every masked pixel is its own target (x = src, y = ref at the same pixel);
        runs entirely in the CUDA kernels.  Rows with a non-finite channel in either image are
        dropped (:35-36) and fewer than 200 remaining samples give the identity (:38-41).
      * ``"ot"`` (the reference's behaviour: Sinkhorn barycentric targets from POT, :47-56) is not
        available: POT is absent, unpinned, and there is no CPU fallback here — NotImplementedError
        rather than a silent substitution.
    """
    if targets == "ot":
        raise NotImplementedError(
            "fit_ot_poly_rgb(targets='ot') needs the Sinkhorn barycentric-target stage "
            "(reference poly_regression.py:47-56, third-party POT) which is not part of this build; "
            "pass targets='paired' for the pixel-paired GPU fit")
    if targets != "paired":
        raise ValueError("targets must be 'ot' or 'paired'")
    numpy_in = is_numpy_like(src_rgb)
    src = to_device(src_rgb, torch.float32)
    ref = to_device(ref_rgb, torch.float32, src.device)
    if src.shape != ref.shape or src.dim() != 3:
        raise ValueError(f"src_rgb / ref_rgb must be (H,W,C) of equal shape, got {tuple(src.shape)}, {tuple(ref.shape)}")
    m = to_device(mask, torch.uint8, src.device)
    xs = src.permute(2, 0, 1).contiguous()
    ys = ref.permute(2, 0, 1).contiguous()
    fm = kernels.fit_mask(xs, m, gate_k=-1)
    fm = kernels.fit_mask(ys, fm, gate_k=-1)
    coeffs = kernels.poly_fit(xs, ys, fm, int(deg), min_count=200)
    return to_host(coeffs) if numpy_in else coeffs


def apply_poly_rgb(rgb, coeffs, mask=None):
    """Apply per-channel polynomial mapping to an (H, W, C) image in [0, 1] (reference :65-84):
    float32 copy, float64 Horner where ``mask`` (everywhere if None), then clip ALL pixels to [0, 1]."""
    numpy_in = is_numpy_like(rgb)
    x = to_device(rgb, torch.float32)
    c = to_device(coeffs, torch.float64, x.device)
    m = None if mask is None else to_device(mask, torch.uint8, x.device)
    out = kernels.poly_apply(x, c, m, lo=0.0, hi=1.0, layout="interleaved")
    return to_host(out, np.float32) if numpy_in else out

"""Per-channel polynomial colour matching — call surface of the reference's
``s2_emit/poly_regression.py`` (``fit_ot_poly_rgb`` :16-62, ``apply_poly_rgb`` :65-84).

The import-time script of the reference (:86-172, hard-coded /content paths) is deliberately not
reproduced.  The least-squares core (np.polyfit at :58-60) runs as fp64 normal-equation moments
+ a warp-level solve on the GPU; the OT target stage (:31-56) as a streamed fp64 Sinkhorn (csrc/ot.cu);
the apply is a fused Horner + mask + clip kernel.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import kernels
from .._host import is_numpy_like, repair_rank_deficient, to_device, to_host


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
    coeffs, mom = kernels.poly_fit(xt, yt, mt, deg, min_count=min_count, return_moments=True)
    if not numpy_in:
        return coeffs
    # host results: a rank-deficient band (fewer distinct x than deg + 1) gets np.polyfit's minimum-norm answer and its
    # RankWarning instead of the device solve's NaN (CUDA-tensor callers keep the NaN: no host round trip there)
    return repair_rank_deficient(to_host(coeffs), to_host(mom), int(deg), min_count)


def fit_ot_poly_rgb(src_rgb, ref_rgb, mask, deg=2, n_samples=5000, reg=0.05, numItermax=300, stopThr=1e-6,
                    seed=0, *, targets="ot", return_info=False):
    """Fit per-channel polynomial mapping y = poly(x) using OT barycentric targets; returns (3, deg+1) float64,
    highest power first.  Signature and behaviour of the reference (:16-62) plus two keyword-only extras.

    ``targets="ot"`` (default, the reference): masked rows with every channel finite (:33-36; src and ref are
    filtered independently), fewer than 200 of either -> identity (:38-41), ``default_rng(seed).choice`` samples
    (:46-47, drawn on the host with numpy's generator, gathered on the GPU), squared-euclidean cost, Sinkhorn
    (reg, numItermax, stopThr), barycentric targets (:52-56) and the per-channel polynomial fit (:58-60) all on
    the GPU in fp64.  POT, which the reference calls for the cost matrix and Sinkhorn, is neither vendored nor
    pinned by it; the kernels follow POT's published ``dist`` / ``sinkhorn_knopp`` (csrc/ot.cu).
    ``targets="paired"``: every masked pixel is its own target (x = src, y = ref at the same pixel; rows with a
    non-finite channel in either image dropped) — the pixel-paired fit, no OT.
    """
    if targets not in ("ot", "paired"):
        raise ValueError("targets must be 'ot' or 'paired'")
    numpy_in = is_numpy_like(src_rgb)
    src = to_device(src_rgb, torch.float32)
    ref = to_device(ref_rgb, torch.float32, src.device)
    if src.shape != ref.shape or src.dim() != 3:
        raise ValueError(f"src_rgb / ref_rgb must be (H,W,C) of equal shape, got {tuple(src.shape)}, {tuple(ref.shape)}")
    m = to_device(mask, torch.uint8, src.device)
    if tuple(m.shape) != tuple(src.shape[:2]):
        raise IndexError(f"boolean index did not match: mask {tuple(m.shape)} vs image {tuple(src.shape[:2])}")
    C = src.shape[2]
    info = None
    mom = None
    if targets == "paired":
        xs = src.permute(2, 0, 1).contiguous()
        ys = ref.permute(2, 0, 1).contiguous()
        fm = kernels.fit_mask(xs, m, gate_k=-1, y=ys)
        coeffs, mom = kernels.poly_fit(xs, ys, fm, int(deg), min_count=200, return_moments=True)
    else:
        x2, y2 = src.reshape(-1, C), ref.reshape(-1, C)
        idx_x, nx = kernels.compact_finite_rows(x2, m.reshape(-1))
        idx_y, ny = kernels.compact_finite_rows(y2, m.reshape(-1))
        nx, ny = int(nx.item()), int(ny.item())                    # the only host round trip: the sample draw needs them
        if nx < 200 or ny < 200:                                   # :38-41
            coeffs = torch.zeros((C, int(deg) + 1), dtype=torch.float64, device=src.device)
            coeffs[:, -2] = 1.0
        else:
            rng = np.random.default_rng(seed)                      # :31
            ns, nt = min(int(n_samples), nx), min(int(n_samples), ny)
            sel_x = torch.from_numpy(rng.choice(nx, size=ns, replace=False).astype(np.int64)).to(src.device)   # :46
            sel_y = torch.from_numpy(rng.choice(ny, size=nt, replace=False).astype(np.int64)).to(src.device)   # :47
            X = kernels.gather_rows_f64(x2, idx_x, sel_x)
            Y = kernels.gather_rows_f64(y2, idx_y, sel_y)
            ybar, info = kernels.sinkhorn_barycentric(X, Y, reg, numItermax, stopThr)                           # :49-56
            coeffs, mom = kernels.polyfit_f64(X, ybar, int(deg), return_moments=True)                          # :58-60
    out = to_host(coeffs) if numpy_in else coeffs
    if numpy_in and mom is not None:                               # np.polyfit's answer for rank-deficient channels
        out = repair_rank_deficient(out, to_host(mom), int(deg), 200 if targets == "paired" else 0)
    if return_info:
        keys = ("iterations", "err", "err_iteration", "numerical_error")
        return out, (None if info is None else dict(zip(keys, to_host(info).tolist())))
    return out


def apply_poly_rgb(rgb, coeffs, mask=None):
    """Apply per-channel polynomial mapping to an (H, W, C) image in [0, 1] (reference :65-84):
    float32 copy, float64 Horner where ``mask`` (everywhere if None), then clip ALL pixels to [0, 1]."""
    numpy_in = is_numpy_like(rgb)
    x = to_device(rgb, torch.float32)
    c = to_device(coeffs, torch.float64, x.device)
    m = None if mask is None else to_device(mask, torch.uint8, x.device)
    out = kernels.poly_apply(x, c, m, lo=0.0, hi=1.0, layout="interleaved")
    return to_host(out, np.float32) if numpy_in else out

"""Oracle for polynomial colour matching (test infrastructure — see oracle/__init__.py).

Restates the least-squares core and the apply step of the reference's
``s2_emit/poly_regression.py`` (:33-41 finite filter + identity fallback, :58-60 np.polyfit per
channel, :65-84 apply_poly_rgb) and the pixel-paired per-band calibration of
``Pairs_EMIT_S2_demo-2.ipynb`` cell 72 (np.polyfit over all valid pixels, float64).
The Sinkhorn/OT target construction (:47-56, third-party POT) is NOT restated (parity unpinned).
"""
from __future__ import annotations

import numpy as np


def fit_mask(x_planes, valid=None, gate_k: int = 0, gate_gt: float = 0.0) -> np.ndarray:
    """valid & isfinite(all bands) & (x[gate_k] > gate_gt)   (poly_regression.py:106)."""
    x = np.asarray(x_planes)
    m = np.isfinite(x).all(axis=0)
    if gate_k >= 0:
        with np.errstate(invalid="ignore"):
            m &= x[gate_k] > gate_gt
    if valid is not None:
        m &= np.asarray(valid).astype(bool)
    return m


def polyfit_paired(x_planes, y_planes, mask, deg: int, min_count: int = 0) -> np.ndarray:
    """(K, deg+1) float64: np.polyfit(x[k][m], y[k][m], deg), m = mask & finite(x[k]) & finite(y[k])."""
    x = np.asarray(x_planes)
    y = np.asarray(y_planes)
    K = x.shape[0]
    coeffs = np.zeros((K, deg + 1), dtype=np.float64)
    for k in range(K):
        m = np.isfinite(x[k]) & np.isfinite(y[k])
        if mask is not None:
            mk = np.asarray(mask).astype(bool)
            m &= mk[k] if mk.ndim == x.ndim else mk
        xs = x[k][m].astype(np.float64)
        ys = y[k][m].astype(np.float64)
        if xs.size < min_count:
            coeffs[k, -2] = 1.0                      # identity, poly_regression.py:38-41
        else:
            coeffs[k] = np.polyfit(xs, ys, deg=deg)  # poly_regression.py:58-60
    return coeffs


def fit_poly_rgb_paired(src_rgb, ref_rgb, mask, deg: int = 2) -> np.ndarray:
    """fit_ot_poly_rgb with every masked pixel as its own target: the finite filter (:33-36), the
    <200-sample identity (:38-41) and the per-channel np.polyfit (:58-60) of the reference,
    without the OT resampling in between."""
    m = np.asarray(mask).astype(bool)
    X = np.asarray(src_rgb)[m].reshape(-1, src_rgb.shape[-1]).astype(np.float64)
    Y = np.asarray(ref_rgb)[m].reshape(-1, ref_rgb.shape[-1]).astype(np.float64)
    keep = np.isfinite(X).all(axis=1) & np.isfinite(Y).all(axis=1)
    X, Y = X[keep], Y[keep]
    C = X.shape[1]
    coeffs = np.zeros((C, deg + 1), dtype=np.float64)
    if X.shape[0] < 200:
        coeffs[:, -2] = 1.0
        return coeffs
    for c in range(C):
        coeffs[c] = np.polyfit(X[:, c], Y[:, c], deg=deg)
    return coeffs


def apply_poly_rgb(rgb, coeffs, mask=None) -> np.ndarray:
    """float32 copy -> polyval (float64) per channel, written where mask -> clip ALL to [0, 1]   (:65-84)."""
    out = np.array(rgb, dtype=np.float32, copy=True)
    C = out.shape[-1]
    for c in range(C):
        x = out[..., c]
        y = np.polyval(coeffs[c], x)
        if mask is None:
            out[..., c] = y
        else:
            m = np.asarray(mask).astype(bool)
            ch = x.copy()
            ch[m] = y[m]
            out[..., c] = ch
    return np.clip(out, 0.0, 1.0)


def apply_poly_planes(x_planes, coeffs, mask=None, lo: float = 0.0, hi: float = 1.0) -> np.ndarray:
    """Planar (K, ...) flavour of apply_poly_rgb."""
    x = np.asarray(x_planes, dtype=np.float32)
    moved = np.moveaxis(x, 0, -1)
    m = None
    if mask is not None:
        m = np.asarray(mask).astype(bool)
        if m.ndim == x.ndim:
            raise ValueError("apply_poly_planes takes one mask shared by all planes")
    out = np.array(moved, dtype=np.float32, copy=True)
    for c in range(out.shape[-1]):
        xc = out[..., c]
        y = np.polyval(coeffs[c], xc)
        if m is None:
            out[..., c] = y
        else:
            ch = xc.copy()
            ch[m] = y[m]
            out[..., c] = ch
    if lo <= hi:
        out = np.clip(out, lo, hi)
    return np.ascontiguousarray(np.moveaxis(out, -1, 0))

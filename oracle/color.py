"""Oracle for the shared percentile stretch (test infrastructure — see oracle/__init__.py).

Restates ``apply_shared_percentile_stretch`` of the reference's ``s2_emit/color.py:25-34`` (the step at
``s2_emit/poly_regression.py:126-127``), plus a from-first-principles restatement of what
``np.percentile(vals, q)`` (method "linear") computes, used to check that the device interpolation follows
numpy operation by operation.
"""
from __future__ import annotations

import numpy as np


def apply_shared_percentile_stretch(img, mask, pmin: float = 2, pmax: float = 98) -> np.ndarray:
    img = np.asarray(img)
    out = np.zeros_like(img, dtype=np.float32)                                   # :29
    for c in range(img.shape[-1]):                                               # :30 (the reference: range(3))
        vals = img[..., c][mask]                                                 # :31
        lo, hi = np.percentile(vals, [pmin, pmax])                               # :32
        out[..., c] = np.clip((img[..., c] - lo) / (hi - lo + 1e-12), 0, 1)      # :33
    return out


def shared_percentile_limits(img, mask, pmin: float = 2, pmax: float = 98) -> np.ndarray:
    """(C, 2) float64 of the (lo, hi) pairs the stretch uses (color.py:31-32)."""
    img = np.asarray(img)
    return np.stack([np.percentile(img[..., c][mask], [pmin, pmax]) for c in range(img.shape[-1])])


def percentile_linear_sorted(vals: np.ndarray, q_percent) -> np.ndarray:
    """np.percentile(vals, q) for a 1-D float32 array, written out: full sort, virtual index (n-1)*q/100,
    float32 neighbour difference, float64 lerp with numpy's gamma >= 0.5 branch, NaN if any NaN."""
    v = np.sort(np.asarray(vals, dtype=np.float32))           # NaNs sort last
    n = v.size
    out = []
    for q in np.atleast_1d(q_percent):
        qf = np.true_divide(np.float64(q), 100)
        vi = (n - 1) * qf
        prev = np.floor(vi)
        if vi >= n - 1:
            a = b = v[-1]
            gamma = vi - (-1.0)
        else:
            a, b = v[int(prev)], v[int(prev) + 1]
            gamma = vi - prev
        diff = np.float32(b) - np.float32(a)                   # float32 subtraction
        r = np.float64(a) + np.float64(diff) * gamma
        if gamma >= 0.5:
            r = np.float64(b) - np.float64(diff) * (1 - gamma)
        if n and np.isnan(v[-1]):
            r = np.float64(np.nan)
        out.append(r)
    return np.array(out, dtype=np.float64)


def ot_match_rgb_sinkhorn_pot(src_rgb, ref_rgb, mask, n_samples=5_000, reg=0.05, numItermax=300, stopThr=1e-6, seed=0):
    """Restatement of the reference's 3-D colour transfer (s2_emit/color.py:63-116).  Its ``ot.dist`` / ``ot.sinkhorn``
    are POT's (absent, unpinned): oracle/ot.py restates them — PARITY UNPINNED for those two calls, pinned for the
    rest by running the reference's own function with oracle/ot.py injected (tests/golden/make_golden_color.py)."""
    from . import ot as oot
    rng = np.random.default_rng(seed)                                            # :78
    X_all = src_rgb[mask].reshape(-1, 3).astype(np.float64)                      # :80-81
    Y_all = ref_rgb[mask].reshape(-1, 3).astype(np.float64)
    X_all = X_all[np.isfinite(X_all).all(axis=1)]                                # :83-84
    Y_all = Y_all[np.isfinite(Y_all).all(axis=1)]
    if X_all.shape[0] < 2 or Y_all.shape[0] < 2:                                 # :86-87
        return src_rgb.copy()
    ns, nt = min(n_samples, X_all.shape[0]), min(n_samples, Y_all.shape[0])
    X = X_all[rng.choice(X_all.shape[0], size=ns, replace=False)]                # :92-93
    Y = Y_all[rng.choice(Y_all.shape[0], size=nt, replace=False)]
    Ybar = oot.barycentric_targets(X, Y, reg, numItermax, stopThr)               # :95-102
    X_aug = np.concatenate([X, np.ones((ns, 1))], axis=1)                        # :104-107
    W, *_ = np.linalg.lstsq(X_aug, Ybar, rcond=None)
    A, t = W[:3, :], W[3, :]
    out = src_rgb.copy().astype(np.float32)                                      # :109-114
    Xm2 = np.clip(out[mask].reshape(-1, 3).astype(np.float64) @ A + t, 0.0, 1.0)
    out[mask] = Xm2.reshape(out[mask].shape).astype(np.float32)
    return out


def robust_norm(x, pmin: float = 2, pmax: float = 98) -> np.ndarray:
    lo, hi = np.nanpercentile(x, [pmin, pmax])                                   # :7
    return np.clip((x - lo) / (hi - lo + 1e-12), 0, 1)                           # :8


def robust_norm_rgb(img, mask, pmin: float = 2, pmax: float = 98) -> np.ndarray:
    y = np.zeros_like(img, dtype=float)                                          # :17
    for c in range(img.shape[-1]):                                               # :18 (the reference: range(3))
        vals = img[..., c][mask]
        lo, hi = np.percentile(vals, [pmin, pmax])                               # :20
        cc = (img[..., c] - lo) / (hi - lo + 1e-12)
        cc[~mask] = np.nan                                                       # :22
        y[..., c] = np.clip(cc, 0, 1)
    return y


def _hist_match_channel(src, ref, mask):
    src_vals = src[mask].ravel()                                                 # :37-38
    ref_vals = ref[mask].ravel()
    s_values, s_idx, s_counts = np.unique(src_vals, return_inverse=True, return_counts=True)   # :40
    r_values, r_counts = np.unique(ref_vals, return_counts=True)                 # :41
    s_quant = np.cumsum(s_counts).astype(np.float64)                             # :43-44
    s_quant /= (s_quant[-1] + 1e-32)
    r_quant = np.cumsum(r_counts).astype(np.float64)
    r_quant /= (r_quant[-1] + 1e-32)
    matched = np.interp(s_quant, r_quant, r_values)[s_idx].reshape(src_vals.shape)   # :46-47
    out = src.copy()                                                             # :49-52
    out[mask] = matched
    return out


def histogram_match_rgb(src_rgb, ref_rgb, mask):
    out = src_rgb.copy()                                                         # :59
    for c in range(src_rgb.shape[-1]):                                           # :60 (the reference: range(3))
        out[..., c] = _hist_match_channel(out[..., c], ref_rgb[..., c], mask)
    return np.clip(out, 0, 1)                                                    # :62

"""Oracle for the optimal-transport target stage of ``fit_ot_poly_rgb`` (test infrastructure — see
oracle/__init__.py).

PARITY UNPINNED for the POT part: the reference calls ``ot.dist`` and ``ot.sinkhorn`` (POT, import name
``ot``; s2_emit/poly_regression.py:4,52-53) but neither vendors nor pins that package, it is not installed in
this image, and the reference holds no test or golden vector at that boundary.  ``dist_sqeuclidean`` and
``sinkhorn_knopp`` below restate POT's PUBLISHED algorithms (ot.utils.euclidean_distances(squared=True) and
ot.bregman.sinkhorn_knopp, the default ``method="sinkhorn"`` of ``ot.sinkhorn``, POT 0.8-0.9 line).  What IS
pinned: everything around them — ``fit_ot_poly_rgb`` itself is the reference's OWN function, AST-loaded from
/root/reference with this module injected as ``ot`` (tests/golden/make_golden.py), so the masking, the finite
filter, the ``default_rng(seed).choice`` sampling, the barycentric projection and ``np.polyfit`` are the
reference's code, not a restatement.
"""
from __future__ import annotations

import warnings

import numpy as np


def dist(x1, x2=None, metric="sqeuclidean"):
    """ot.dist for the one metric the reference uses."""
    if metric != "sqeuclidean":
        raise NotImplementedError("only metric='sqeuclidean' is restated (the reference uses no other)")
    return dist_sqeuclidean(x1, x1 if x2 is None else x2)


def dist_sqeuclidean(X, Y):
    """POT euclidean_distances(X, Y, squared=True): |x|^2 + |y|^2 - 2 x.y, clamped at 0."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    a2 = np.einsum("ij,ij->i", X, X)
    b2 = np.einsum("ij,ij->i", Y, Y)
    c = -2 * np.dot(X, Y.T)
    c += a2[:, None]
    c += b2[None, :]
    c = np.maximum(c, 0)
    if X is Y:
        c = c * (1 - np.eye(X.shape[0], dtype=c.dtype))
    return c


def sinkhorn_knopp(a, b, M, reg, numItermax=1000, stopThr=1e-9, log=False, warn=False):
    """POT ot.bregman.sinkhorn_knopp for 1-D b: returns the transport plan diag(u) K diag(v)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    M = np.asarray(M, dtype=np.float64)
    dim_a, dim_b = len(a), b.shape[0]
    u = np.ones(dim_a, dtype=np.float64) / dim_a
    v = np.ones(dim_b, dtype=np.float64) / dim_b
    K = np.exp(M / (-reg))
    Kp = (1 / a).reshape(-1, 1) * K
    err = 1.0
    info = {"err": [], "niter": 0, "numerical": False}
    ii = -1
    for ii in range(numItermax):
        uprev, vprev = u, v
        KtransposeU = np.dot(K.T, u)
        v = b / KtransposeU
        u = 1.0 / np.dot(Kp, v)
        if (np.any(KtransposeU == 0) or np.any(np.isnan(u)) or np.any(np.isnan(v))
                or np.any(np.isinf(u)) or np.any(np.isinf(v))):
            if warn:
                warnings.warn("Warning: numerical errors at iteration %d" % ii)
            u, v = uprev, vprev
            info["numerical"] = True
            break
        if ii % 10 == 0:
            tmp2 = np.einsum("i,ij,j->j", u, K, v)
            err = np.linalg.norm(tmp2 - b)
            info["err"].append(err)
            if err < stopThr:
                break
    info["niter"] = ii
    P = u.reshape((-1, 1)) * K * v.reshape((1, -1))
    return (P, info) if log else P


def sinkhorn(a, b, M, reg, method="sinkhorn", numItermax=1000, stopThr=1e-9, **kwargs):
    """ot.sinkhorn with its default method."""
    if method.lower() != "sinkhorn":
        raise NotImplementedError("only method='sinkhorn' (sinkhorn_knopp) is restated")
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        return sinkhorn_knopp(a, b, M, reg, numItermax=numItermax, stopThr=stopThr)


def barycentric_targets(X, Y, reg=0.05, numItermax=300, stopThr=1e-6):
    """poly_regression.py:49-56 on already sampled X (ns, C), Y (nt, C) float64."""
    ns, nt = X.shape[0], Y.shape[0]
    a = np.full(ns, 1.0 / ns, dtype=np.float64)
    b = np.full(nt, 1.0 / nt, dtype=np.float64)
    M = dist(X, Y, metric="sqeuclidean")
    P = sinkhorn(a, b, M, reg=reg, numItermax=numItermax, stopThr=stopThr)
    row_sum = P.sum(axis=1, keepdims=True) + 1e-32
    return (P @ Y) / row_sum


def fit_ot_poly_rgb(src_rgb, ref_rgb, mask, deg=2, n_samples=5000, reg=0.05, numItermax=300, stopThr=1e-6, seed=0):
    """Restatement of s2_emit/poly_regression.py:16-62 (the golden vectors come from the reference's own
    function; this copy lets the GPU tests run where /root/reference is absent)."""
    rng = np.random.default_rng(seed)                                             # :31
    C = src_rgb.shape[-1]
    X_all = src_rgb[mask].reshape(-1, C).astype(np.float64)                       # :33
    Y_all = ref_rgb[mask].reshape(-1, C).astype(np.float64)                       # :34
    X_all = X_all[np.isfinite(X_all).all(axis=1)]                                 # :35
    Y_all = Y_all[np.isfinite(Y_all).all(axis=1)]                                 # :36
    if X_all.shape[0] < 200 or Y_all.shape[0] < 200:                              # :38-41
        coeffs = np.zeros((C, deg + 1), dtype=np.float64)
        coeffs[:, -2] = 1.0
        return coeffs
    ns = min(n_samples, X_all.shape[0])                                           # :43-44
    nt = min(n_samples, Y_all.shape[0])
    X = X_all[rng.choice(X_all.shape[0], size=ns, replace=False)]                 # :46-47
    Y = Y_all[rng.choice(Y_all.shape[0], size=nt, replace=False)]
    Ybar = barycentric_targets(X, Y, reg, numItermax, stopThr)                    # :49-56
    coeffs = np.zeros((C, deg + 1), dtype=np.float64)
    for c in range(C):                                                            # :58-60
        coeffs[c] = np.polyfit(X[:, c], Ybar[:, c], deg=deg)
    return coeffs

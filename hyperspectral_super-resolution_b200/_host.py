"""Host <-> device marshalling shared by the reference-compatible wrappers.

The reference's call surface is numpy-in / numpy-out; these helpers move such arrays to the
current CUDA device, and back, so the arithmetic always runs in the CUDA kernels.  torch CUDA
tensors pass through untouched (and results stay on the device).
"""
from __future__ import annotations

import numpy as np
import torch


def cuda_device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("hsr_b200 needs a CUDA device (B200, sm_100a): there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def is_numpy_like(a) -> bool:
    return not isinstance(a, torch.Tensor)


def to_device(a, dtype: torch.dtype, device=None) -> torch.Tensor:
    """numpy array / sequence / tensor -> CUDA tensor of ``dtype`` (bool arrays become uint8 0/1)."""
    if isinstance(a, torch.Tensor):
        t = a
        if not t.is_cuda:
            t = t.to(device or cuda_device(), non_blocking=True)
    else:
        arr = np.asarray(a)
        if arr.dtype == np.bool_:
            arr = arr.view(np.uint8)
        if not arr.flags.c_contiguous:
            arr = np.ascontiguousarray(arr)
        if not arr.flags.writeable:
            arr = arr.copy()
        t = torch.from_numpy(arr).to(device or cuda_device(), non_blocking=True)
    if t.dtype == torch.bool and dtype == torch.uint8:
        t = t.view(torch.uint8)
    elif t.dtype != dtype:
        t = t.to(dtype)
    return t


def to_host(t: torch.Tensor, np_dtype=None) -> np.ndarray:
    a = t.detach().cpu().numpy()
    if np_dtype is not None and a.dtype != np_dtype:
        a = a.astype(np_dtype)
    return a


# ------------------------------------------------------------------ rank-deficient polynomial fits (host results only)
RANK_CUT = 1.4e-14      # eigenvalues of the scaled normal matrix below RANK_CUT * largest count as zero (64 eps: its noise floor)


def min_norm_from_moments(mom, deg: int):
    """What ``np.polyfit`` returns for a RANK-DEFICIENT series (fewer distinct x than deg + 1, e.g. a constant band),
    from the series' normal-equation moments ``[S_0 .. S_2deg, T_0 .. T_deg]`` (S_j = sum x^j, T_j = sum x^j y).

    np.polyfit (s2_emit/poly_regression.py:58-60) solves the column-scaled Vandermonde system by SVD least squares and
    so returns the minimum-norm solution of the scaled system; the device solve is a Gauss-Jordan elimination of the
    same scaled normal equations and yields NaN when a pivot vanishes.  Here the scaled normal matrix is diagonalised
    (its eigenvectors are the right singular vectors of the scaled Vandermonde matrix, its eigenvalues the squared
    singular values), null directions are dropped and the rest solved — the same answer.  Returns ``(coeffs, rank)``,
    coefficients highest power first.  Only a repair for host-side (numpy) results: the normal matrix resolves singular
    values down to ~1e-7 of the largest, np.polyfit's cut-off is len(x) * eps."""
    mom = np.asarray(mom, dtype=np.float64)
    N = int(deg) + 1
    S, T = mom[:2 * deg + 1], mom[2 * deg + 1:]
    scale = np.sqrt(np.maximum(S[0:2 * deg + 1:2], 0.0))
    scale = np.where(scale > 0.0, scale, 1.0)
    A = np.array([[S[i + j] / (scale[i] * scale[j]) for j in range(N)] for i in range(N)])
    b = T / scale
    lam, Q = np.linalg.eigh(A)
    keep = lam > RANK_CUT * max(lam.max(), 0.0)
    z = Q[:, keep] @ ((Q[:, keep].T @ b) / lam[keep])
    return (z / scale)[::-1].copy(), int(keep.sum())


def repair_rank_deficient(coeffs: np.ndarray, moments: np.ndarray, deg: int, min_count: int = 0) -> np.ndarray:
    """Replace the coefficient rows of rank-deficient series (see ``min_norm_from_moments``) and warn as numpy does."""
    import warnings

    out = np.array(coeffs, dtype=np.float64, copy=True)
    flat, mom = out.reshape(-1, int(deg) + 1), np.asarray(moments, dtype=np.float64).reshape(-1, 3 * int(deg) + 2)
    for s in range(flat.shape[0]):
        if not (mom[s, 0] >= max(int(min_count), 1)) or not np.isfinite(mom[s]).all():
            continue                                    # identity fallback / non-finite inputs: the device's answer stands
        c, rank = min_norm_from_moments(mom[s], deg)
        if rank < int(deg) + 1:
            flat[s] = c
            warnings.warn("Polyfit may be poorly conditioned", np.exceptions.RankWarning, stacklevel=3)
    return out

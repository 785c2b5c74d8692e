"""Synthetic EMIT-granule-shaped inputs: there is no network for real granules.

Everything is generated from explicit seeds.  The numpy flavours produce the small cases the CPU
oracle checks; the torch flavours build the full-size benchmark inputs directly on the device.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np
import torch

EMIT_BANDS = 285
GRANULE_RAW_SHAPE = (1280, 1242, EMIT_BANDS)      # (downtrack, crosstrack, bands)
MASKED_BAND_VALUE = -0.01                          # EMIT L2A value of masked bands


def emit_wavelengths(bands: int = EMIT_BANDS) -> np.ndarray:
    """float32 (bands,) grid starting at 381.00558 nm, ~7.4 nm spacing (Spectral_matching.ipynb cell 16)."""
    return np.linspace(381.00558, 2492.92, bands).astype(np.float32)


def good_band_mask(emit_w: np.ndarray) -> np.ndarray:
    """EMIT ``good_wavelengths``: the water-vapour windows 1320-1440 nm and 1770-1970 nm are masked."""
    w = np.asarray(emit_w, dtype=np.float64)
    return ~(((w > 1320) & (w < 1440)) | ((w > 1770) & (w < 1970)))


def rotation_glt(raw_h: int, raw_w: int, theta_deg: float = 25.0) -> Tuple[np.ndarray, np.ndarray]:
    """Nearest-neighbour rotation GLT about the centre: (glt_x, glt_y) int32 [Ho, Wo], 1-based, 0 = nodata.

    Ho = ceil(Hr cos + Wr sin), Wo = ceil(Wr cos + Hr sin); for 1280 x 1242 at 25 deg this is
    1685 x 1667 with 56.6 % valid entries.
    """
    th = math.radians(theta_deg)
    c, s = math.cos(th), math.sin(th)
    Ho = int(math.ceil(raw_h * c + raw_w * s))
    Wo = int(math.ceil(raw_w * c + raw_h * s))
    yy, xx = np.meshgrid(np.arange(Ho, dtype=np.float64), np.arange(Wo, dtype=np.float64), indexing="ij")
    cx, cy = (Wo - 1) / 2.0, (Ho - 1) / 2.0
    rx = np.rint((xx - cx) * c + (yy - cy) * s + (raw_w - 1) / 2.0)
    ry = np.rint(-(xx - cx) * s + (yy - cy) * c + (raw_h - 1) / 2.0)
    ok = (rx >= 0) & (rx < raw_w) & (ry >= 0) & (ry < raw_h)
    glt_x = np.where(ok, rx + 1, 0).astype(np.int32)
    glt_y = np.where(ok, ry + 1, 0).astype(np.int32)
    return glt_x, glt_y


def inject_glt_defects(glt_x: np.ndarray, glt_y: np.ndarray, raw_h: int, raw_w: int, seed: int = 7,
                       hole_frac: float = 0.001, n_oob: int = 64, n_neg: int = 64):
    """Parity-run defects: random holes (0), out-of-range entries (gx = Wr + 5) and negative entries."""
    rng = np.random.default_rng(seed)
    gx, gy = glt_x.copy(), glt_y.copy()
    n = gx.size
    flat_x, flat_y = gx.reshape(-1), gy.reshape(-1)
    holes = rng.choice(n, size=max(1, int(n * hole_frac)), replace=False)
    flat_x[holes] = 0
    oob = rng.choice(n, size=min(n_oob, n), replace=False)
    flat_x[oob] = raw_w + 5
    flat_y[oob] = np.maximum(flat_y[oob], 1)
    neg = rng.choice(n, size=min(n_neg, n), replace=False)
    flat_y[neg] = -3
    flat_x[neg] = np.maximum(flat_x[neg], 1)
    return gx, gy


def identity_glt(h: int, w: int, zero_frac: float = 0.02, seed: int = 2) -> Tuple[np.ndarray, np.ndarray]:
    """Tile GLT (BASELINE config 3): identity + 1 with a fraction of random nodata entries."""
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.arange(h, dtype=np.int32), np.arange(w, dtype=np.int32), indexing="ij")
    gx, gy = xx + 1, yy + 1
    holes = rng.random((h, w)) < zero_frac
    gx = np.where(holes, 0, gx).astype(np.int32)
    gy = np.where(holes, 0, gy).astype(np.int32)
    return gx, gy


# ----------------------------------------------------------------------------------- numpy cubes
def raw_cube_bits_np(shape, seed: int = 0, good: Optional[np.ndarray] = None) -> np.ndarray:
    """Bit-exactness flavour: every element a distinct-looking finite fp32 (0.6 * U[0,1))."""
    rng = np.random.default_rng(seed)
    raw = (0.6 * rng.random(shape, dtype=np.float32)).astype(np.float32)
    if good is not None:
        raw[..., ~np.asarray(good)] = MASKED_BAND_VALUE
    return raw


def raw_cube_spectra_np(shape, seed: int = 0, good: Optional[np.ndarray] = None) -> np.ndarray:
    """Bench / polyfit flavour: smooth spectra, x spans ~[0.03, 0.8]  (SURVEY.md 8d-1)."""
    rng = np.random.default_rng(seed)
    h, w, nb = shape
    a = rng.uniform(0.05, 0.8, size=(h, w, 1)).astype(np.float32)
    phi = rng.uniform(0.0, 2 * np.pi, size=(h, w, 1)).astype(np.float32)
    b = np.arange(nb, dtype=np.float32)[None, None, :]
    raw = a * (0.6 + 0.4 * np.sin(0.02 * b + phi)) + 0.02 * (rng.random(shape, dtype=np.float32) - 0.5)
    raw = np.clip(raw, 0.0, 1.0).astype(np.float32)
    if good is not None:
        raw[..., ~np.asarray(good)] = MASKED_BAND_VALUE
    return raw


def s2_reference_np(x_planes: np.ndarray, seed: int = 1, noise: float = 0.005) -> np.ndarray:
    """Synthetic 'real S2' planes on the same grid: y_k = c2_k x^2 + c1 x + c0 + N(0, noise)."""
    rng = np.random.default_rng(seed)
    K = x_planes.shape[0]
    y = np.empty_like(x_planes, dtype=np.float32)
    for k in range(K):
        c2, c1, c0 = -0.3 + 0.02 * k, 1.1, 0.02
        xk = x_planes[k].astype(np.float64)
        y[k] = (c2 * xk * xk + c1 * xk + c0 + rng.normal(0.0, noise, size=xk.shape)).astype(np.float32)
    return y


# ----------------------------------------------------------------------------------- torch cubes
def raw_cube_spectra_torch(shape, seed: int, device, good: Optional[np.ndarray] = None) -> torch.Tensor:
    """Device-side generator of the bench flavour (same formula as ``raw_cube_spectra_np``, torch RNG)."""
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    h, w, nb = shape
    a = 0.05 + 0.75 * torch.rand((h, w, 1), generator=g, device=device)
    phi = (2 * math.pi) * torch.rand((h, w, 1), generator=g, device=device)
    b = torch.arange(nb, device=device, dtype=torch.float32)[None, None, :]
    raw = torch.empty(shape, dtype=torch.float32, device=device)
    rows = max(1, (64 << 20) // (w * nb * 4))          # build in ~64 MB row slabs
    for r0 in range(0, h, rows):
        r1 = min(h, r0 + rows)
        blk = a[r0:r1] * (0.6 + 0.4 * torch.sin(0.02 * b + phi[r0:r1]))
        blk += 0.02 * (torch.rand((r1 - r0, w, nb), generator=g, device=device) - 0.5)
        raw[r0:r1] = blk.clamp_(0.0, 1.0)
    if good is not None:
        bad = torch.from_numpy(~np.asarray(good)).to(device)
        raw[..., bad] = MASKED_BAND_VALUE
    return raw


def s2_reference_torch(x_planes: torch.Tensor, seed: int = 1, noise: float = 0.005) -> torch.Tensor:
    g = torch.Generator(device=x_planes.device)
    g.manual_seed(int(seed))
    K = x_planes.shape[0]
    k = torch.arange(K, device=x_planes.device, dtype=torch.float32).view(K, *([1] * (x_planes.dim() - 1)))
    c2 = -0.3 + 0.02 * k
    x = torch.nan_to_num(x_planes, nan=0.0, posinf=0.0, neginf=0.0)
    y = c2 * x * x + 1.1 * x + 0.02
    y += noise * torch.randn(x_planes.shape, generator=g, device=x_planes.device)
    return y.float()

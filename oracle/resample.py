"""Oracle for the aligned "average" downsampling (test infrastructure — see oracle/__init__.py).

PARITY UNPINNED against GDAL: the reference calls ``rasterio.warp.reproject(..., Resampling.average)``
(Pairs_EMIT_S2_demo-2.ipynb cell 73), neither rasterio nor GDAL is installed here, and the reference holds no
vector at that boundary.  Restated for the one geometry its pipeline produces — same CRS, grids snapped with an
integer pixel ratio (emit_proj.py:794-797) — where GDAL's average is the mean of the valid source pixels under each
destination pixel (accumulated in double, cast to the destination type), 0 where none is valid.
"""
from __future__ import annotations

import numpy as np


def downsample_to_grid(src_stack, factor=6, src_scale=None, nodata=None):
    s = np.asarray(src_stack)
    C, Hs, Ws = s.shape
    Hd, Wd = Hs // factor, Ws // factor
    blk = s[:, :Hd * factor, :Wd * factor].astype(np.float64).reshape(C, Hd, factor, Wd, factor)
    use = ~np.isnan(blk)
    if nodata is not None:
        use &= blk != float(nodata)
    out = np.zeros((C, Hd, Wd), dtype=np.float32)                      # destination initialised to 0 (cell 73)
    # sequential row-major accumulation inside each block, as a per-pixel loop does
    acc = np.zeros((C, Hd, Wd), dtype=np.float64)
    cnt = np.zeros((C, Hd, Wd), dtype=np.int64)
    for r in range(factor):
        for q in range(factor):
            v, u = blk[:, :, r, :, q], use[:, :, r, :, q]
            acc += np.where(u, v, 0.0)
            cnt += u
    nz = cnt > 0
    out[nz] = (acc[nz] / cnt[nz]).astype(np.float32)
    if src_scale is not None:
        out *= float(src_scale)                                        # cell 73: out *= float(src_scale)
    return out


def upsample_to_grid(src_stack, factor=6, nodata=None):
    """Bilinear onto the factor-times finer aligned grid (GDAL's 4-sample bilinear kernel restated; parity unpinned):
    2 x 2 neighbours of the destination pixel centre, unusable neighbours skipped, weights renormalised, else 0."""
    s = np.asarray(src_stack, dtype=np.float32)
    C, Hs, Ws = s.shape
    Hd, Wd = Hs * factor, Ws * factor
    sy = (np.arange(Hd, dtype=np.float64) + 0.5) / factor - 0.5
    sx = (np.arange(Wd, dtype=np.float64) + 0.5) / factor - 0.5
    y0, x0 = np.floor(sy).astype(np.int64), np.floor(sx).astype(np.int64)
    wy1, wx1 = sy - np.floor(sy), sx - np.floor(sx)
    acc = np.zeros((C, Hd, Wd), dtype=np.float64)
    wsum = np.zeros((C, Hd, Wd), dtype=np.float64)
    for dy in (0, 1):
        yy = y0 + dy
        wy = wy1 if dy else 1.0 - wy1
        oky = (yy >= 0) & (yy < Hs)
        for dx in (0, 1):
            xx = x0 + dx
            wx = wx1 if dx else 1.0 - wx1
            okx = (xx >= 0) & (xx < Ws)
            w = wy[:, None] * wx[None, :]
            v = s[:, np.clip(yy, 0, Hs - 1)][:, :, np.clip(xx, 0, Ws - 1)].astype(np.float64)
            use = (oky[:, None] & okx[None, :] & (w != 0.0))[None] & ~np.isnan(v)
            if nodata is not None:
                use &= v != np.float32(nodata)
            acc += np.where(use, w[None] * v, 0.0)
            wsum += np.where(use, w[None], 0.0)
    out = np.zeros((C, Hd, Wd), dtype=np.float32)
    nz = wsum > 0
    out[nz] = (acc[nz] / wsum[nz]).astype(np.float32)
    return out

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

"""Oracle for GLT orthorectification (test infrastructure — see oracle/__init__.py).

Restates, in numpy:
  * the production rule of ``nc_to_envi``:   EMIT_data/emit_proj.py:646-661 (raw dim order),
    :682-703 (GLT -> int32, validity, 0-based, in-bounds), :705-718 (diagnostics),
    :947-948 (index lists), :968-987 (chunked fill + gather), :1123-1131 / :1217-1224 (planes);
  * the xarray-flavour ``apply_glt``:        EMIT_data/emit_tools.py:153-181.
"""
from __future__ import annotations

import numpy as np

NO_DATA_VALUE = -9999.0  # emit_proj.py:27


def glt_to_int32(glt_x, glt_y) -> np.ndarray:
    """[Ho, Wo, 2] int32 (x, y) with NaN -> 0   (emit_proj.py:683-687)."""
    gx = np.asarray(glt_x)
    gy = np.asarray(glt_y)
    glt = np.empty(gx.shape + (2,), dtype=np.int32)
    glt[..., 0] = np.nan_to_num(gx, nan=0).astype(np.int32)
    glt[..., 1] = np.nan_to_num(gy, nan=0).astype(np.int32)
    return glt


def glt_validity(glt: np.ndarray, raw_h: int, raw_w: int):
    """(glt0, valid, valid_inbounds, diag)   (emit_proj.py:691-703, :713-718)."""
    valid = (glt[..., 0] != 0) & (glt[..., 1] != 0)              # :691  all(glt != 0, axis=-1)
    glt0 = glt.copy()
    glt0[valid] -= 1                                             # :693-694 (int32 arithmetic, wraps like numpy)
    x0 = glt0[..., 0]
    y0 = glt0[..., 1]
    inb = (y0 >= 0) & (y0 < raw_h) & (x0 >= 0) & (x0 < raw_w)    # :698-701
    valid2 = valid & inb                                         # :703
    n_valid = int(np.count_nonzero(valid))
    n_inb = int(np.count_nonzero(valid2))
    diag = {
        "raw_shape_yx": [int(raw_h), int(raw_w)],
        "valid_glt_count": n_valid,
        "valid_glt_inbounds_count": n_inb,
        "valid_glt_dropped_oob": n_valid - n_inb,
    }
    return glt0, valid, valid2, diag


def glt_ortho(raw, glt_x, glt_y, fill: float = NO_DATA_VALUE, transpose_raw_yx: bool = False, chunk: int = 32):
    """(ortho [Ho, Wo, B] f32, valid_inbounds bool, diag) following nc_to_envi's data export."""
    raw = np.asarray(raw)
    glt = glt_to_int32(glt_x, glt_y)
    d0, d1, nb = raw.shape
    raw_h, raw_w = (d1, d0) if transpose_raw_yx else (d0, d1)    # :696
    glt0, _, valid2, diag = glt_validity(glt, raw_h, raw_w)
    gy = glt0[..., 1][valid2]                                    # :947
    gx = glt0[..., 0][valid2]                                    # :948
    Ho, Wo = glt.shape[:2]
    out = np.empty((Ho, Wo, nb), dtype=np.float32)
    for b0 in range(0, nb, chunk):                               # :971
        b1 = min(b0 + chunk, nb)
        blk = np.asarray(raw[:, :, b0:b1], dtype=np.float32)     # :975
        if transpose_raw_yx:
            blk = blk.transpose(1, 0, 2)                         # :976-977
        tile = np.full((Ho, Wo, b1 - b0), fill, dtype=np.float32)   # :981
        tile[valid2, :] = blk[gy, gx, :]                         # :982
        out[:, :, b0:b1] = tile
    return out, valid2, diag


def glt_plane(band, glt_x, glt_y, fill: float = NO_DATA_VALUE, transpose_raw_yx: bool = False):
    """One 2-D plane onto the GLT grid (emit_proj.py:1123-1131, :1217-1224)."""
    band = np.asarray(band, dtype=np.float32)
    if transpose_raw_yx:
        band = band.T
    glt = glt_to_int32(glt_x, glt_y)
    glt0, _, valid2, _ = glt_validity(glt, band.shape[0], band.shape[1])
    out = np.full(glt.shape[:2], fill, dtype=np.float32)
    out[valid2] = band[glt0[..., 1][valid2], glt0[..., 0][valid2]]
    return out


def apply_glt(ds_array, glt_array, fill_value=-9999, nodata=0):
    """emit_tools.py:153-181: no bounds test — negative entries wrap, too-large entries raise."""
    arr = np.asarray(ds_array)
    if arr.ndim == 2:
        arr = arr[:, :, None]                                    # :166-167
    glt = np.asarray(glt_array)
    out = np.full(glt.shape[:2] + (arr.shape[-1],), fill_value, dtype=np.float32)   # :168-172
    ok = np.all(glt != nodata, axis=-1)                          # :173
    idx = glt.copy()                                             # :176
    idx[ok] -= 1                                                 # :177
    out[ok, :] = arr[idx[ok, 1], idx[ok, 0], :]                  # :178-180
    return out

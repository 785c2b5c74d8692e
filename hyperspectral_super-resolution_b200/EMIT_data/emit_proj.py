"""GLT orthorectification — array-level core of the reference's ``EMIT_data/emit_proj.py``.

The reference's ``nc_to_envi`` (:563-1300) interleaves file I/O (netCDF in, ENVI/GeoTIFF out,
``gdalwarp`` subprocesses) with the arithmetic this module replaces:

    GLT assembly + validity + in-bounds rule      emit_proj.py:682-703
    gather index lists                            emit_proj.py:947-948
    32-band chunked fill + gather                 emit_proj.py:968-987
    LOC / OBS plane gathers                       emit_proj.py:1123-1131, :1217-1224

``glt_ortho`` / ``ortho_planes`` are that arithmetic as CUDA kernels; ``nc_to_envi`` and
``convert_emit_nc_to_envi`` (``EMIT_data/nc_export.py``, re-exported here) keep the reference's signatures and
drive them from files when the optional I/O dependencies (netCDF4 or h5netcdf; GDAL CLI + rasterio for the UTM
warp) are installed.
"""
from __future__ import annotations

from typing import Dict, Sequence

import numpy as np
import torch

from .. import kernels
from .._host import is_numpy_like, to_device, to_host

NO_DATA_VALUE = -9999.0   # reference emit_proj.py:27


def _diag_dict(diag: torch.Tensor, raw_h: int, raw_w: int) -> Dict[str, object]:
    v = [int(t) for t in diag.cpu().tolist()]
    return {  # keys of info["glt_diag"], emit_proj.py:713-718
        "raw_shape_yx": [int(raw_h), int(raw_w)],
        "valid_glt_count": v[0],
        "valid_glt_inbounds_count": v[1],
        "valid_glt_dropped_oob": v[2],
    }


def glt_ortho(raw, glt_x, glt_y, *, fill: float = NO_DATA_VALUE, transpose_raw_yx: bool = False):
    """Orthorectify ``raw`` [Hr, Wr, B] (or [Wr, Hr, B] with ``transpose_raw_yx``, emit_proj.py:646-661)
    with the 1-based GLT planes ``glt_x`` / ``glt_y`` [Ho, Wo] (float with NaN, or integer).

    Returns ``(ortho [Ho, Wo, B] float32, valid_mask [Ho, Wo] bool, diag dict)``; numpy in -> numpy out,
    CUDA tensors in -> CUDA tensors out.  Bit-exact with ``out[valid] = raw[gy, gx, :]`` over a
    ``-9999`` filled cube.
    """
    numpy_in = is_numpy_like(raw)
    r = to_device(raw, torch.float32)
    gx, gy = kernels.prepare_glt(glt_x, glt_y, device=r.device)
    ortho, valid, diag = kernels.glt_ortho(r, gx, gy, fill=float(fill), transpose_raw_yx=transpose_raw_yx)
    d0, d1 = r.shape[0], r.shape[1]
    raw_h, raw_w = (d1, d0) if transpose_raw_yx else (d0, d1)
    info = _diag_dict(diag, raw_h, raw_w)
    if numpy_in:
        return to_host(ortho, np.float32), to_host(valid).astype(bool), info
    return ortho, valid, info


def ortho_planes(planes: Sequence, glt_x, glt_y, *, fill: float = NO_DATA_VALUE, transpose_raw_yx: bool = False):
    """Gather 2-D raw-space planes (lon / lat / elev, OBS bands) onto the GLT grid
    (emit_proj.py:1123-1131, :1217-1224).  ``planes`` is a sequence of [Hr, Wr] arrays (``.T`` of
    the file layout is taken when ``transpose_raw_yx``); returns a list of [Ho, Wo] float32 planes."""
    if len(planes) == 0:
        return []
    numpy_in = is_numpy_like(planes[0])
    dev = to_device(planes[0], torch.float32).device
    gx, gy = kernels.prepare_glt(glt_x, glt_y, device=dev)
    outs = []
    for pl in planes:
        t = to_device(pl, torch.float32, dev)
        o, _, _ = kernels.glt_ortho(t, gx, gy, fill=float(fill), transpose_raw_yx=transpose_raw_yx,
                                    want_valid=False, want_diag=False)
        outs.append(to_host(o, np.float32) if numpy_in else o)
    return outs


def __getattr__(name):   # file-level drivers live in nc_export.py (lazy: they pull json / subprocess / pathlib only)
    if name in ("nc_to_envi", "convert_emit_nc_to_envi", "get_attr", "open_any_nc", "run_cmd", "write_envi_bil"):
        from . import nc_export
        return getattr(nc_export, name)
    if name in ("export_uint16_deflate_geotiff", "raster_meta", "export_loc_uint16_deflate_geotiff",
                "export_obs_uint16_deflate_geotiff", "_sample_band_minmax", "_which_gdal_edit"):
        from . import gdal_export                       # GDAL-facing file plumbing (reference :248-306, :392-560)
        return getattr(gdal_export, name)
    if name in ("_compute_te", "_intersect", "_bounds_to_out_crs"):
        from . import warp                              # extent arithmetic of the UTM step (reference :309-382)
        return getattr(warp, {"_compute_te": "compute_te"}.get(name, name))
    raise AttributeError(name)

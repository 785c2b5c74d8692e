"""uint16 GeoTIFF exports and raster metadata — the GDAL-facing helpers of the reference's ``EMIT_data/emit_proj.py``
(``export_uint16_deflate_geotiff`` :248-277, ``raster_meta`` :281-306, ``export_loc_uint16_deflate_geotiff`` :399-455,
``_sample_band_minmax`` :458-493, ``export_obs_uint16_deflate_geotiff`` :495-560), kept importable with the reference's
signatures.

They are file-format plumbing around OTHER PROGRAMS (``gdal_translate`` / ``gdal_edit`` subprocesses, rasterio readers),
not arithmetic of the hot path: each function assembles the command line the reference issues and hands it to
``run_cmd``; the returned records (including the ``uint16_decode`` scale / offset tables) have the reference's keys.
Nothing here touches the GPU, and nothing is emulated: without the GDAL command-line tools the subprocess fails
(``FileNotFoundError``), without rasterio the functions that read rasters raise ``ImportError``.
"""
from __future__ import annotations

import shutil
from pathlib import Path
from typing import Optional, Tuple

import numpy as np

from .nc_export import run_cmd

_GTIFF_U16 = ["gdal_translate", "-of", "GTiff", "-ot", "UInt16"]


def _which_gdal_edit() -> Optional[str]:
    """``gdal_edit`` or ``gdal_edit.py``, whichever is on PATH (reference :392-396)."""
    return next((c for c in ("gdal_edit", "gdal_edit.py") if shutil.which(c)), None)


def _rasterio():
    try:
        import rasterio
    except ImportError as e:
        raise ImportError("this helper reads rasters with rasterio, which is not installed") from e
    return rasterio


def export_uint16_deflate_geotiff(src_path: str, dst_tif: str, *, assign_epsg: Optional[str] = None,
                                  scale_mode: str = "none", nodata_uint16: int = 65535, zlevel: int = 1) -> dict:
    """DEFLATE-compressed UInt16 GeoTIFF of a raster; ``scale_mode="emit_reflectance_0_1"`` maps [0, 1] to
    [0, 10000] and records the decode metadata (reference :248-277)."""
    cmd = _GTIFF_U16 + ["-co", "COMPRESS=DEFLATE", "-co", f"ZLEVEL={int(zlevel)}", "-co", "PREDICTOR=2",
                        "-co", "NUM_THREADS=ALL_CPUS", "-co", "BIGTIFF=IF_SAFER"]
    if scale_mode == "emit_reflectance_0_1":
        nd = str(int(nodata_uint16))
        cmd += ["-scale", "0", "1", "0", "10000", "-a_nodata", nd, "-mo", "scale_factor=0.0001", "-mo", "units=reflectance",
                "-mo", f"uint16_nodata={nd}"]
    if assign_epsg:
        cmd += ["-a_srs", assign_epsg]
    return run_cmd(cmd + [src_path, dst_tif], check=True)


def raster_meta(path: str) -> dict:
    """CRS / bounds / shape / resolution of any GDAL-readable raster (reference :281-306)."""
    p = Path(path)
    if not p.exists():
        return {"path": str(path), "exists": False}
    rasterio = _rasterio()
    from rasterio.warp import transform_bounds  # pragma: no cover  (rasterio is not installable in the build image)
    with rasterio.open(str(p)) as ds:  # pragma: no cover
        b, crs = ds.bounds, ds.crs
        out = {"path": str(p), "exists": True, "driver": ds.driver, "crs": crs.to_string() if crs else None,
               "width": ds.width, "height": ds.height, "count": ds.count,
               "res": [float(ds.res[0]), float(ds.res[1])] if ds.res else None,
               "bounds": [float(b.left), float(b.bottom), float(b.right), float(b.top)], "nodata": ds.nodata}
        if crs:
            out["bounds_wgs84"] = list(transform_bounds(crs, "EPSG:4326", b.left, b.bottom, b.right, b.top, densify_pts=21))
        return out


def _band_scale_args(ranges) -> list:
    """``-scale_<b> lo hi 0 65535 -exponent_<b> 1`` per band (the exponent makes gdal_translate clamp to the range)."""
    args = []
    for b, (lo, hi) in enumerate(ranges, start=1):
        args += [f"-scale_{b}", str(lo), str(hi), "0", "65535", f"-exponent_{b}", "1"]
    return args


def _decode_record(rec: dict, dst_tif: str, mins, maxs, nodata_uint16: int, extra: dict) -> dict:
    """Attach (and, if gdal_edit exists, write) the per-band decode table: true = raw * scale + offset."""
    scales = [(hi - lo) / 65535.0 for lo, hi in zip(mins, maxs)]
    offsets = list(mins)
    tool = _which_gdal_edit()
    if tool:
        run_cmd([tool, "-scale", *[f"{s:.16g}" for s in scales], "-offset", *[f"{o:.16g}" for o in offsets], dst_tif],
                check=True)
    rec["uint16_decode"] = {"scales": scales, "offsets": offsets, **extra, "nodata_uint16": nodata_uint16,
                            "note": "Recover: true = raw*scale + offset"}
    return rec


def export_loc_uint16_deflate_geotiff(src_path: str, dst_tif: str, *, lon_range=(-180.0, 180.0), lat_range=(-90.0, 90.0),
                                      elev_range=(-1000.0, 12000.0), nodata_uint16: int = 0) -> dict:
    """EMIT LOC (lon, lat, elev) as a UInt16 GeoTIFF, each band scaled over its physical range (reference :399-455)."""
    ranges = [tuple(lon_range), tuple(lat_range), tuple(elev_range)]
    cmd = _GTIFF_U16 + ["-a_nodata", str(nodata_uint16), "-co", "COMPRESS=DEFLATE", "-co", "PREDICTOR=2", "-co", "TILED=YES"]
    rec = run_cmd(cmd + _band_scale_args(ranges) + [src_path, dst_tif], check=True)
    return _decode_record(rec, dst_tif, [r[0] for r in ranges], [r[1] for r in ranges], nodata_uint16,
                          {"ranges": [list(r) for r in ranges]})


def _sample_band_minmax(src_path: str, band_index_1based: int, nodata: float, *, stride: int = 64, p_low: float = 1.0,
                        p_high: float = 99.0) -> Tuple[float, float]:
    """Robust (percentile) range of one band from a ``stride``-decimated nearest-neighbour read (reference :458-493)."""
    rasterio = _rasterio()
    from rasterio.enums import Resampling  # pragma: no cover
    with rasterio.open(src_path) as ds:  # pragma: no cover
        shape = (max(1, ds.height // stride), max(1, ds.width // stride))
        arr = ds.read(band_index_1based, out_shape=shape, resampling=Resampling.nearest).astype(np.float32)
    return _robust_range(arr, nodata, p_low, p_high)  # pragma: no cover


def _robust_range(arr: np.ndarray, nodata: float, p_low: float, p_high: float) -> Tuple[float, float]:
    """The range rule of ``_sample_band_minmax``: percentiles of the finite, non-nodata samples; min / max when they
    collapse; (0, 1) when nothing is valid; never a zero-width range."""
    vals = arr[np.isfinite(arr) & (arr != float(nodata))]
    if vals.size == 0:
        return 0.0, 1.0
    lo, hi = np.percentile(vals, [p_low, p_high])
    if not np.isfinite(lo) or not np.isfinite(hi) or lo == hi:
        lo, hi = float(vals.min()), float(vals.max())
        if lo == hi:
            hi = lo + 1.0
    return float(lo), float(hi)


def export_obs_uint16_deflate_geotiff(src_path: str, dst_tif: str, *, nodata_float: float, nodata_uint16: int = 0,
                                      stride: int = 64, p_low: float = 1.0, p_high: float = 99.0) -> dict:
    """EMIT OBS cube as a UInt16 GeoTIFF with per-band robust scaling (reference :495-560)."""
    rasterio = _rasterio()
    with rasterio.open(src_path) as ds:  # pragma: no cover
        nb = ds.count
    ranges = [_sample_band_minmax(src_path, b, nodata_float, stride=stride, p_low=p_low, p_high=p_high)  # pragma: no cover
              for b in range(1, nb + 1)]
    cmd = _GTIFF_U16 + ["-a_nodata", str(nodata_uint16), "-co", "COMPRESS=DEFLATE", "-co", "PREDICTOR=2", "-co", "TILED=YES"]  # pragma: no cover
    rec = run_cmd(cmd + _band_scale_args(ranges) + [src_path, dst_tif], check=True)  # pragma: no cover
    mins, maxs = [r[0] for r in ranges], [r[1] for r in ranges]  # pragma: no cover
    return _decode_record(rec, dst_tif, mins, maxs, nodata_uint16,  # pragma: no cover
                          {"src_mins": mins, "src_maxs": maxs, "percentiles": [p_low, p_high], "stride": stride})

"""Sentinel-2 stack -> EMIT grid, the aligned integer-ratio case of the notebook's ``downsample_s2_to_grid``
(Pairs_EMIT_S2_demo-2.ipynb cell 73; called at ``s2_emit/poly_regression.py:110-116``).

The reference reprojects with ``rasterio.warp.reproject(..., resampling=average)``; ``nc_to_envi`` snaps the EMIT grid
to the Sentinel-2 origin with an integer pixel ratio (``emit_proj.py:794-797``), so the warp degenerates to the mean of
every ``factor x factor`` block, which is what runs here (on the GPU, from an in-memory stack instead of file paths).
Everything else goes through the general warp kernel (``csrc/warp.cu``); parity with GDAL is unpinned.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import kernels
from .._host import cuda_device, is_numpy_like, to_host


def downsample_to_grid(src_stack, factor: int = 6, src_scale=None, nodata=None):
    """(nbands, Hs, Ws) uint8 / uint16 / float32 -> (nbands, Hs // factor, Ws // factor) float32: block mean
    ("average" resampling onto the aligned coarser grid), then ``* src_scale`` as the reference does."""
    numpy_in = is_numpy_like(src_stack)
    if numpy_in:
        a = np.ascontiguousarray(src_stack)
        if a.dtype == np.uint16:                 # torch.from_numpy has no uint16 on older builds: go through int16 bits
            t = torch.from_numpy(a.view(np.int16)).to(cuda_device()).view(torch.uint16)
        elif a.dtype in (np.uint8, np.float32):
            t = torch.from_numpy(a).to(cuda_device())
        else:
            t = torch.from_numpy(a.astype(np.float32)).to(cuda_device())
    else:
        t = src_stack
    out = kernels.block_average(t, factor, nodata=nodata, scale=src_scale)
    return to_host(out, np.float32) if numpy_in else out


def upsample_to_grid(src_stack, factor: int = 6, nodata=None):
    """(C, Hs, Ws) float32 -> (C, Hs*factor, Ws*factor) float32, bilinear, aligned grids — the aligned case of the
    notebook's ``reproject_stack_to_grid(..., "bilinear")`` (pseudo-S2 planes 60 m -> 10 m, poly_regression.py:150-156)."""
    numpy_in = is_numpy_like(src_stack)
    t = torch.from_numpy(np.ascontiguousarray(src_stack, dtype=np.float32)).to(cuda_device()) if numpy_in else src_stack
    out = kernels.bilinear_upsample(t, factor, nodata=nodata)
    return to_host(out, np.float32) if numpy_in else out


# ------------------------------------------------------------------ the notebook's names, grids instead of file paths
def _grid(g):
    from ..EMIT_data.warp import S2Grid
    return S2Grid.coerce(g)            # an S2Grid, a dict (epsg, x0, y0, dx, dy, width, height) or a raster path (rasterio)


def _gt(g):
    return (g.x0, g.dx, 0.0, g.y0, 0.0, -g.dy)


def _aligned_factor(fine, coarse):
    """k if every pixel of ``coarse`` is exactly k x k pixels of ``fine`` (same CRS, same origin, whole extent), else 0."""
    if fine.epsg != coarse.epsg:
        return 0
    k = coarse.dx / fine.dx
    ok = (abs(k - round(k)) < 1e-9 and abs(coarse.dy / fine.dy - k) < 1e-9 and abs(fine.x0 - coarse.x0) < 1e-6 * fine.dx
          and abs(fine.y0 - coarse.y0) < 1e-6 * fine.dy and coarse.width * round(k) == fine.width
          and coarse.height * round(k) == fine.height)
    return int(round(k)) if ok and round(k) >= 1 else 0


def downsample_s2_to_grid(src_stack, src_grid, dst_grid, band_indexes=None, src_scale=None, resampling="average"):
    """The notebook's ``downsample_s2_to_grid`` (Pairs_EMIT_S2_demo-2.ipynb cell 73; ``s2_emit/poly_regression.py:110-116``)
    on an in-memory (C, Hs, Ws) stack: the 1-based ``band_indexes`` of it onto ``dst_grid``, float32, ``* src_scale``.
    ``average`` on the snapped geometry nc_to_envi produces (same CRS and origin, integer pixel ratio) is the block mean
    kernel; other geometries / kernels go through ``reproject_stack_to_grid`` (the general warp)."""
    sg, dg = _grid(src_grid), _grid(dst_grid)
    stack = src_stack if band_indexes is None else src_stack[[int(b) - 1 for b in band_indexes]]
    if resampling == "average":
        k = _aligned_factor(sg, dg)
        if k:
            return downsample_to_grid(stack, k, src_scale=src_scale)
    out = reproject_stack_to_grid(stack, sg, dg, resampling)
    if src_scale is not None:
        out = out * np.float32(src_scale) if is_numpy_like(out) else out.mul_(float(src_scale))
    return out


def reproject_stack_to_grid(src_stack, src_grid, dst_grid, resampling="bilinear"):
    """The notebook's ``reproject_stack_to_grid`` (cell 73; ``s2_emit/poly_regression.py:150-156``): a (C, H, W) float32
    stack from ``src_grid`` onto ``dst_grid`` -> (C, H2, W2) float32, 0 where the destination is not covered.
    ``bilinear`` onto a finer snapped grid is the aligned kernel; anything else (shifted / rotated grids, another UTM zone
    is NOT supported: same CRS, or geographic WGS-84 -> UTM) is the general warp kernel (``nearest`` / ``bilinear`` /
    ``cubic`` / ``average`` — the resamplings s2_data/s2_utils.py:546-574 and the notebook ask rasterio for)."""
    from ..EMIT_data.warp import warp_to_grid
    sg, dg = _grid(src_grid), _grid(dst_grid)
    if resampling not in ("nearest", "bilinear", "cubic", "average"):
        raise NotImplementedError(f"resampling {resampling!r}: 'nearest', 'bilinear', 'cubic' and 'average' are implemented")
    k = _aligned_factor(dg, sg)
    if resampling == "bilinear" and k:
        return upsample_to_grid(src_stack, k)
    if sg.epsg == dg.epsg:
        epsg = None
    elif sg.epsg == 4326:
        epsg = dg.epsg
    else:
        raise NotImplementedError(f"EPSG:{sg.epsg} -> EPSG:{dg.epsg}: same CRS or EPSG:4326 -> UTM only")
    numpy_in = is_numpy_like(src_stack)
    t = torch.from_numpy(np.ascontiguousarray(src_stack, dtype=np.float32)).to(cuda_device()) if numpy_in else src_stack
    cube = t.permute(1, 2, 0).contiguous()                                     # planes -> band-interleaved records
    out = warp_to_grid(cube, _gt(sg), _gt(dg), (dg.height, dg.width), epsg=epsg, kernel=resampling, nodata=None)
    out = out.permute(2, 0, 1).contiguous()
    return to_host(out, np.float32) if numpy_in else out

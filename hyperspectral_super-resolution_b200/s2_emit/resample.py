"""Sentinel-2 stack -> EMIT grid, the aligned integer-ratio case of the notebook's ``downsample_s2_to_grid``
(Pairs_EMIT_S2_demo-2.ipynb cell 73; called at ``s2_emit/poly_regression.py:110-116``).

The reference reprojects with ``rasterio.warp.reproject(..., resampling=average)``; ``nc_to_envi`` snaps the EMIT grid
to the Sentinel-2 origin with an integer pixel ratio (``emit_proj.py:794-797``), so the warp degenerates to the mean of
every ``factor x factor`` block, which is what runs here (on the GPU, from an in-memory stack instead of file paths).
Other CRSs, rotations and the cubic / bilinear kernels of GDAL are out of scope; parity with GDAL is unpinned.
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

"""Tile validity, quantisation and band subsampling — the array arithmetic of the reference's
``tiles_helpers/utils.py`` (``is_black_mask`` :201-220, the window walk of ``find_valid_paired_tiles`` :223-305,
the uint16 quantisation inside ``save_tile_pair`` :362-373, ``_subsample_bands_evenly`` :444-458).

numpy in -> numpy out, CUDA tensors in -> CUDA tensors out; the arithmetic always runs on the GPU.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import kernels
from .._host import is_numpy_like, to_device, to_host


def is_black_mask(arr, nodata=None, masked_val=-0.01, nodata_atol=1e-3, zero_atol=1e-6):
    """arr: (bands, H, W) float32.  Pixel is black/invalid if all bands ~ nodata, or all ~ masked_val, or all ~ 0
    (reference :201-220).  Returns (H, W) bool."""
    numpy_in = is_numpy_like(arr)
    a = to_device(arr, torch.float32)
    if a.dim() != 3:
        raise ValueError(f"arr must be (bands, H, W), got {tuple(a.shape)}")
    m = kernels.black_mask(a, nodata, masked_val, nodata_atol, zero_atol)
    return to_host(m) if numpy_in else m


def quantize_emit_u16(emit, nodata=None, emit_scale=10000.0, emit_nodata_u16=65535):
    """The EMIT branch of save_tile_pair (reference :357-371): reflectance -> uint16 ``rint(x * emit_scale)`` clipped to
    [0, emit_nodata_u16 - 1]; non-finite samples and samples equal to ``nodata`` become ``emit_nodata_u16``."""
    numpy_in = is_numpy_like(emit)
    a = to_device(emit, torch.float32)
    q = kernels.quantize_u16(a, nodata, emit_scale, emit_nodata_u16)
    return to_host(q.view(torch.int16)).view(np.uint16) if numpy_in else q


def _subsample_bands_evenly(num_bands_total, num_keep=32):
    """Evenly spaced band indices in [0, num_bands_total - 1] (reference :444-458): rounded linspace, de-duplicated,
    then midpoints of consecutive picks until ``num_keep`` indices exist.  Host-side integer logic."""
    idx = np.unique(np.linspace(0, num_bands_total - 1, num_keep).round().astype(int))
    while len(idx) < num_keep:
        missing = num_keep - len(idx)
        mids = [int((idx[i] + idx[i + 1]) // 2) for i in range(min(missing, len(idx) - 1))]
        idx = np.unique(np.concatenate([idx, np.array(mids, dtype=int)]))
    return idx[:num_keep]


def subsample_bands(arr, num_keep=32, idx_0based=None):
    """(bands, H, W) -> (num_keep, H, W) with the band picks of write_emit_b32_tile (reference :460-491).
    Returns ``(subset, idx_0based)``."""
    nb = arr.shape[0]
    if idx_0based is None:
        if nb < num_keep:
            raise ValueError(f"Tile has only {nb} bands, can't keep {num_keep}.")
        idx_0based = _subsample_bands_evenly(nb, num_keep=num_keep)
    idx_0based = np.asarray(idx_0based, dtype=int)
    if is_numpy_like(arr):
        return np.asarray(arr)[idx_0based], idx_0based
    return arr[torch.as_tensor(idx_0based, device=arr.device)], idx_0based


def find_valid_paired_tiles_arrays(emit, s2, emit_tile_size=100, scale=6, max_black_frac=0.0, max_tiles=None,
                                   emit_nodata=None, s2_nodata=None):
    """find_valid_paired_tiles (reference :223-305) on in-memory rasters instead of file paths.

    emit: (bands, He, We) float32; s2: (bands, Hs, Ws) float32 on a ``scale``-times finer grid.  Walks the
    non-overlapping ``emit_tile_size`` windows in the reference's order and keeps those whose black fraction
    (is_black_mask) is <= max_black_frac in BOTH rasters.  Returns a list of dicts with ``idx``,
    ``emit_window`` / ``s2_window`` as (col_off, row_off, width, height) and the two black fractions.
    The masks are computed once per raster, the per-window sums by one block per window.
    """
    e = to_device(emit, torch.float32)
    s = to_device(s2, torch.float32, e.device)
    if e.dim() != 3 or s.dim() != 3:
        raise ValueError("emit and s2 must be (bands, H, W)")
    h_e, w_e = e.shape[1:]
    h_s, w_s = s.shape[1:]
    te, ts = int(emit_tile_size), int(emit_tile_size) * int(scale)
    e_cnt = to_host(kernels.tile_sums(kernels.black_mask(e, emit_nodata), te, te))
    s_cnt = to_host(kernels.tile_sums(kernels.black_mask(s, s2_nodata), ts, ts))
    tiles, idx = [], 0
    for ty, row_e in enumerate(range(0, h_e - te + 1, te)):
        for tx, col_e in enumerate(range(0, w_e - te + 1, te)):
            row_s, col_s = row_e * scale, col_e * scale
            if (row_s + ts > h_s) or (col_s + ts > w_s):
                continue
            emit_black_frac = e_cnt[ty, tx] / (te * te)
            s2_black_frac = s_cnt[ty, tx] / (ts * ts)
            if emit_black_frac <= max_black_frac and s2_black_frac <= max_black_frac:
                tiles.append({"idx": idx, "emit_window": (col_e, row_e, te, te), "s2_window": (col_s, row_s, ts, ts),
                              "emit_black_frac": emit_black_frac, "s2_black_frac": s2_black_frac})
                idx += 1
                if max_tiles is not None and len(tiles) >= max_tiles:
                    return tiles
    return tiles

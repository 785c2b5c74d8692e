"""Oracle for tile validity / quantisation / band subsampling (test infrastructure — see oracle/__init__.py).

Restates ``tiles_helpers/utils.py``: ``is_black_mask`` :201-220 and ``_subsample_bands_evenly`` :444-458 (both
pinned against the reference's own functions by tests/golden/make_golden.py), the window walk of
``find_valid_paired_tiles`` :258-305 and the uint16 quantisation lines of ``save_tile_pair`` :357-371 (inline in
file-I/O functions, not callable: restated line by line, pinned only through is_black_mask / numpy itself).
"""
from __future__ import annotations

import numpy as np


def is_black_mask(arr, nodata=None, masked_val=-0.01, nodata_atol=1e-3, zero_atol=1e-6):
    if nodata is not None:                                                                     # :211-214
        nodata_mask = np.all(np.isclose(arr, nodata, atol=nodata_atol), axis=0)
    else:
        nodata_mask = np.zeros(arr.shape[1:], dtype=bool)
    masked_mask = np.all(np.isclose(arr, masked_val, atol=nodata_atol), axis=0)               # :216
    zero_mask = np.all(np.abs(arr) < zero_atol, axis=0)                                        # :218
    return nodata_mask | masked_mask | zero_mask                                               # :220


def quantize_emit_u16(emit_tile, nodata=None, emit_scale=10000.0, emit_nodata_u16=65535):
    emit = emit_tile.astype(np.float32, copy=False)                                            # :357
    valid = np.isfinite(emit)                                                                  # :359
    if nodata is not None:                                                                     # :360-361
        valid &= (emit != nodata)
    with np.errstate(invalid="ignore", over="ignore"):
        scaled_i32 = np.rint(emit * float(emit_scale)).astype(np.int32, copy=False)            # :363
    scaled_i32 = np.clip(scaled_i32, 0, int(emit_nodata_u16) - 1)                              # :365
    emit_u16 = np.full(emit.shape, int(emit_nodata_u16), dtype=np.uint16)                      # :367
    emit_u16[valid] = scaled_i32[valid].astype(np.uint16, copy=False)                          # :368
    return emit_u16


def subsample_bands_evenly(num_bands_total, num_keep=32):
    idx = np.linspace(0, num_bands_total - 1, num_keep).round().astype(int)                    # :446
    idx = np.unique(idx)                                                                       # :447
    while len(idx) < num_keep:                                                                 # :449-457
        missing = num_keep - len(idx)
        add = []
        for i in range(len(idx) - 1):
            if len(add) >= missing:
                break
            add.append(int((idx[i] + idx[i + 1]) // 2))
        idx = np.unique(np.concatenate([idx, np.array(add, dtype=int)]))
    return idx[:num_keep]                                                                      # :458


def find_valid_paired_tiles_arrays(emit, s2, emit_tile_size=100, scale=6, max_black_frac=0.0, max_tiles=None,
                                   emit_nodata=None, s2_nodata=None):
    """The loop of find_valid_paired_tiles (:258-305) with array slicing in place of rasterio windows."""
    tiles = []
    h_e, w_e = emit.shape[1:]
    h_s, w_s = s2.shape[1:]
    te = emit_tile_size
    ts = te * scale
    idx = 0
    for row_e in range(0, h_e - te + 1, te):
        for col_e in range(0, w_e - te + 1, te):
            row_s, col_s = row_e * scale, col_e * scale
            if (row_s + ts > h_s) or (col_s + ts > w_s):
                continue
            emit_black = is_black_mask(emit[:, row_e:row_e + te, col_e:col_e + te], nodata=emit_nodata)
            s2_black = is_black_mask(s2[:, row_s:row_s + ts, col_s:col_s + ts], nodata=s2_nodata)
            ef = emit_black.sum() / emit_black.size
            sf = s2_black.sum() / s2_black.size
            if ef <= max_black_frac and sf <= max_black_frac:
                tiles.append({"idx": idx, "emit_window": (col_e, row_e, te, te), "s2_window": (col_s, row_s, ts, ts),
                              "emit_black_frac": ef, "s2_black_frac": sf})
                idx += 1
                if max_tiles is not None and len(tiles) >= max_tiles:
                    return tiles
    return tiles

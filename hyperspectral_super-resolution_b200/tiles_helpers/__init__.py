"""Array-level mirror of the reference's ``tiles_helpers`` (tile validity, uint16 quantisation, band subsample).

The reference's functions are file-path based (rasterio in, GeoTIFF out) — I/O that is out of scope here; what
they COMPUTE on the arrays in between runs in the CUDA kernels of ``csrc/tiles.cu``.
"""
from .utils import (_subsample_bands_evenly, find_valid_paired_tiles_arrays, is_black_mask,  # noqa: F401
                    quantize_emit_u16, subsample_bands)

__all__ = ["is_black_mask", "find_valid_paired_tiles_arrays", "quantize_emit_u16", "subsample_bands",
           "_subsample_bands_evenly"]

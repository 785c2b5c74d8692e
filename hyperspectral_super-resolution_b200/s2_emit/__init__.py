"""Mirror of the reference's ``s2_emit`` call surface for the pair-synthesis hot path.

Unlike the reference's ``s2_emit/__init__.py:1-8`` this does not import h5py / spectral /
rasterio / POT at import time; the SRF, input-format, synthesis and colour-matching names are provided
(plotting and co-registration are outside the hot path).
"""
from .srf import (DEFAULT_SRF_XLSX_URL, S2_BANDS_13, load_s2_srf_from_xlsx, pick_sheet_name,  # noqa: F401
                  srf_fold_weights, synthetic_s2_srf)
from .emit_io import load_emit_envi_rfl, load_emit_wavelengths_from_nc  # noqa: F401
from .synth import pseudo_s2_rgb, pseudo_s2_srf_integral  # noqa: F401
from .poly_regression import apply_poly_rgb, fit_ot_poly_rgb, poly_fit  # noqa: F401
from .color import (apply_shared_percentile_stretch, histogram_match_rgb, ot_match_rgb_sinkhorn_pot, robust_norm,  # noqa: F401
                    robust_norm_rgb,
                    shared_percentile_limits)
from .resample import downsample_s2_to_grid, downsample_to_grid, reproject_stack_to_grid, upsample_to_grid  # noqa: F401
from .pair_matching import match_pair_rgb  # noqa: F401

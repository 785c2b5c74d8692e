"""``apply_glt`` — call surface of the reference's ``EMIT_data/emit_tools.py:153-181``.

The numpy fancy-index gather becomes the CUDA GLT gather kernel.  Only the array function is
mirrored; ``emit_xarray`` / ``ortho_xr`` (:34-125, :184-268) are xarray plumbing around it.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import kernels
from .._host import is_numpy_like, to_device, to_host


def apply_glt(ds_array, glt_array, fill_value=-9999, GLT_NODATA_VALUE=0):
    """Orthorectify a 2-D or 3-D raw-space array with a GLT array ``[Ho, Wo, 2]`` (x, y; 1-based).

    Bit-identical to the reference on GLTs whose non-zero entries are in range.  Entries that are
    negative or beyond the raw grid — where the reference wraps around or raises IndexError
    (emit_tools.py:176-180) — are treated as nodata, the rule of ``nc_to_envi``
    (emit_proj.py:696-703).
    """
    if GLT_NODATA_VALUE != 0:
        raise ValueError("only GLT_NODATA_VALUE = 0 (the EMIT convention) is supported")
    numpy_in = is_numpy_like(ds_array)
    raw = to_device(ds_array, torch.float32)
    if raw.dim() not in (2, 3):
        raise ValueError(f"ds_array must be 2-D or 3-D, got shape {tuple(raw.shape)}")
    glt = glt_array if isinstance(glt_array, torch.Tensor) else np.asarray(glt_array)
    if glt.ndim != 3 or glt.shape[-1] != 2:
        raise ValueError(f"glt_array must be [Ho, Wo, 2], got shape {tuple(glt.shape)}")
    gx, gy = kernels.prepare_glt(glt[..., 0], glt[..., 1], device=raw.device)
    r3 = raw if raw.dim() == 3 else raw.unsqueeze(-1)        # emit_tools.py:166-167
    ortho, _, _ = kernels.glt_ortho(r3, gx, gy, fill=float(fill_value), want_valid=False, want_diag=False)
    return to_host(ortho, np.float32) if numpy_in else ortho

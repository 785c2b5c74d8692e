"""Oracle for SRF band synthesis (test infrastructure — see oracle/__init__.py).

Restates ``pseudo_s2_srf_integral`` and ``pseudo_s2_rgb`` of the reference's
``s2_emit/synth.py:9-58`` in numpy float64, evaluated in row slabs so that granule-sized cubes
do not need the reference's four cube-sized float64 temporaries (same arithmetic per pixel).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

_trapz = getattr(np, "trapezoid", None) or np.trapz


def band_response_on_grid(emit_w: np.ndarray, lam_srf, rsp_srf, good_mask=None) -> np.ndarray:
    """SRF resampled on the EMIT grid, masked bands zeroed   (synth.py:33-35)."""
    rsp = np.interp(emit_w, lam_srf, rsp_srf, left=0.0, right=0.0)
    if good_mask is not None:
        rsp = rsp * np.asarray(good_mask).astype(float)
    return rsp


def pseudo_s2_srf_integral(R, emit_w, srf_dict: Dict[str, Tuple[np.ndarray, np.ndarray]], good_mask=None,
                           rows_per_slab: int = 64) -> Dict[str, Optional[np.ndarray]]:
    """band -> (H, W) float64, or None when the response vanishes on the grid   (synth.py:9-45)."""
    R = np.asarray(R)
    lam = np.asarray(emit_w).astype(float)                                       # :25
    if R.ndim != 3:
        raise ValueError(f"R must be (H,W,B). Got shape {R.shape}")              # :27-28
    if lam.ndim != 1 or lam.shape[0] != R.shape[-1]:
        raise ValueError(f"emit_w must be (B,) matching R bands. Got {lam.shape} vs {R.shape[-1]}")   # :29-30
    H, Wd = R.shape[:2]
    out: Dict[str, Optional[np.ndarray]] = {}
    for band, (lam_srf, rsp_srf) in srf_dict.items():                            # :32
        rsp = band_response_on_grid(lam, lam_srf, rsp_srf, good_mask)
        if np.all(rsp == 0):                                                     # :37-39
            out[band] = None
            continue
        den = _trapz(rsp, x=lam)                                                 # :42
        plane = np.empty((H, Wd), dtype=np.float64)
        for r0 in range(0, H, rows_per_slab):
            slab = R[r0:r0 + rows_per_slab]
            num = _trapz(slab * rsp[None, None, :], x=lam, axis=-1)              # :41
            plane[r0:r0 + rows_per_slab] = num / (den + 1e-32)                   # :43
        out[band] = plane
    return out


def pseudo_s2_rgb(pseudo_s2, order=("B4", "B3", "B2")) -> np.ndarray:
    """(H, W, 3) stack; raises when a band is None / missing   (synth.py:47-58)."""
    planes = []
    for b in order:
        p = pseudo_s2.get(b)
        if p is None:
            raise ValueError(f"Band {b} is None/missing in pseudo_s2.")
        planes.append(p)
    return np.stack(planes, axis=-1)

"""SRF-weighted band synthesis — call surface of the reference's ``s2_emit/synth.py``.

``pseudo_s2_srf_integral`` (reference :9-45) and ``pseudo_s2_rgb`` (:47-58) keep their
signatures; the 13 numpy passes with float64 cube-sized temporaries (synth.py:32-43) become one
pass of the CUDA SRF kernel over host-folded weights (``srf_fold_weights``).
``crop_to_overlap`` (:61-139) is raster file I/O and is outside the hot path.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch

from .. import kernels
from .._host import is_numpy_like, to_device, to_host
from .srf import srf_fold_weights


def pseudo_s2_srf_integral(R, emit_w, srf_dict: Dict[str, Tuple[np.ndarray, np.ndarray]],
                           good_mask=None) -> Dict[str, Optional[np.ndarray]]:
    """band -> (H, W) SRF-weighted mean of the spectrum, ``None`` for bands with no response
    on the (masked) EMIT grid.

    R is (H, W, B): a numpy array (result: float64 numpy, as the reference returns) or a CUDA
    float32 tensor (result: float32 CUDA tensors, no host copy).  Arithmetic is fp32 FMA over
    float64-folded weights: within 1e-5 relative of the reference's float64 path.
    """
    numpy_in = is_numpy_like(R)
    if R.ndim != 3:
        raise ValueError(f"R must be (H,W,B). Got shape {tuple(R.shape)}")
    lam = emit_w.detach().cpu().numpy() if isinstance(emit_w, torch.Tensor) else np.asarray(emit_w)
    if lam.ndim != 1 or lam.shape[0] != R.shape[-1]:
        raise ValueError(f"emit_w must be (B,) matching R bands. Got {lam.shape} vs {R.shape[-1]}")
    good = None
    if good_mask is not None:
        good = good_mask.detach().cpu().numpy() if isinstance(good_mask, torch.Tensor) else np.asarray(good_mask)
    W, names, none_bands, _ = srf_fold_weights(lam, srf_dict, good)

    out: Dict[str, Optional[np.ndarray]] = {}
    planes = None
    if names:
        cube = to_device(R, torch.float32)
        planes = kernels.srf_integrate(cube, to_device(W, torch.float32, cube.device))
        if numpy_in:
            planes = to_host(planes, np.float64)
    col = {b: i for i, b in enumerate(names)}
    for band in srf_dict:                       # keep the reference's dict order
        out[band] = None if band in none_bands else planes[col[band]]
    return out


def pseudo_s2_rgb(pseudo_s2: Dict[str, Optional[np.ndarray]], order=("B4", "B3", "B2")):
    """(H, W, 3) stack of three synthesised bands; raises if one is None/missing (reference :47-58)."""
    chans = []
    for b in order:
        x = pseudo_s2.get(b, None)
        if x is None:
            raise ValueError(f"Band {b} is None/missing in pseudo_s2.")
        chans.append(x)
    if isinstance(chans[0], torch.Tensor):
        return torch.stack(chans, dim=-1)
    return np.stack(chans, axis=-1)

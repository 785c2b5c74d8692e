"""The reference's pair-synthesis SCRIPT as a function: ``s2_emit/poly_regression.py:86-172`` runs at import time on
hard-coded ``/content/...`` files; this is the same sequence of calls on in-memory arrays, every stage on the GPU.

    :97-104   pseudo_s2_srf_integral  -> B2, B3, B4 planes (60 m)                    SRF kernel
    :106,118  valid60 = finite(emit) & emit[B2] > 0 & finite(s2)                       fit-mask kernel
    :110-116  downsample_s2_to_grid(..., "average") * (1 / 255)                        block-average kernel
    :121-127  RGB stacks, apply_shared_percentile_stretch of both images              radix-select percentiles + stretch
    :129-137  fit_ot_poly_rgb(deg = 4, 5000 samples, reg = 0.05, seed = 0)             Sinkhorn + fp64 polyfit
    :139      apply_poly_rgb                                                           Horner + mask + clip
    :150-162  reproject_stack_to_grid(..., "bilinear") to 10 m, stretch, apply again   bilinear kernel, ...

The two resampling calls are the aligned integer-ratio case (grids snapped by nc_to_envi); plotting is dropped.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import kernels
from .._host import to_device, to_host
from . import color, poly_regression, resample, synth


def match_pair_rgb(R, emit_w, srf_dict, good_mask, s2_rgb_10m, *, factor: int = 6, src_scale=1.0 / 255.0, deg: int = 4,
                   n_samples: int = 5000, reg: float = 0.05, numItermax: int = 300, stopThr: float = 1e-6, seed: int = 0,
                   pmin: float = 2, pmax: float = 98, bands_rgb=("B2", "B3", "B4"), to_numpy: bool = True):
    """R (H, W, bands) EMIT reflectance on the 60 m grid, s2_rgb_10m (3, H*factor, W*factor) Sentinel-2 R, G, B
    (uint8 / uint16 / float32).  Returns a dict with ``emit_sim_60m`` (3, H, W: B2, B3, B4), ``valid60``,
    ``s2_real_60m``, ``emit_rgb_n``, ``s2_rgb_n``, ``coeffs`` (3, deg+1), ``matched_60m``, ``emit_sim_10m``,
    ``mask10``, ``matched_10m`` — the variables of the reference's script, numpy arrays unless ``to_numpy=False``."""
    cube = to_device(R, torch.float32)
    pseudo = synth.pseudo_s2_srf_integral(cube, emit_w, srf_dict, good_mask)                       # :102
    missing = [b for b in bands_rgb if pseudo.get(b) is None]
    if missing:
        raise ValueError(f"Band {missing[0]} is None/missing in pseudo_s2.")
    H, W = cube.shape[:2]
    emit_sim = kernels.alloc_planes(3, (H, W), cube.device)
    emit_sim.copy_(torch.stack([pseudo[b] for b in bands_rgb]))                                    # :104 (B, G, R)
    s2_60 = resample.downsample_to_grid(_as_cuda_stack(s2_rgb_10m, cube.device), factor, src_scale=src_scale)   # :110-116
    valid60 = kernels.fit_mask(emit_sim, None, gate_k=0, gate_gt=0.0, y=s2_60)                     # :106, :118
    emit_rgb = emit_sim[[2, 1, 0]].permute(1, 2, 0).contiguous()                                   # :122 (R, G, B)
    s2_rgb = s2_60.permute(1, 2, 0).contiguous()                                                   # :124
    emit_rgb_n = color.apply_shared_percentile_stretch(emit_rgb, valid60, pmin, pmax)              # :126
    s2_rgb_n = color.apply_shared_percentile_stretch(s2_rgb, valid60, pmin, pmax)                  # :127
    coeffs = poly_regression.fit_ot_poly_rgb(emit_rgb_n, s2_rgb_n, valid60, deg=deg, n_samples=n_samples, reg=reg,
                                             numItermax=numItermax, stopThr=stopThr, seed=seed)    # :129-137
    matched60 = poly_regression.apply_poly_rgb(emit_rgb_n, coeffs, mask=valid60)                   # :139
    emit_sim_10 = resample.upsample_to_grid(emit_sim, factor)                                      # :150-155
    emit_rgb_10 = emit_sim_10[[2, 1, 0]].permute(1, 2, 0).contiguous()                             # :157
    mask10 = kernels.fit_mask(emit_sim_10, None, gate_k=-1)                                        # :159
    emit_rgb_10_n = color.apply_shared_percentile_stretch(emit_rgb_10, mask10, pmin, pmax)         # :161
    matched10 = poly_regression.apply_poly_rgb(emit_rgb_10_n, coeffs, mask=mask10)                 # :162
    out = {"emit_sim_60m": emit_sim, "valid60": valid60, "s2_real_60m": s2_60, "emit_rgb_n": emit_rgb_n,
           "s2_rgb_n": s2_rgb_n, "coeffs": coeffs, "matched_60m": matched60, "emit_sim_10m": emit_sim_10,
           "mask10": mask10, "matched_10m": matched10}
    if to_numpy:
        out = {k: to_host(v.contiguous()) for k, v in out.items()}
    return out


def _as_cuda_stack(a, device):
    if isinstance(a, torch.Tensor):
        return a.to(device)
    arr = np.ascontiguousarray(a)
    if arr.dtype == np.uint16:
        return torch.from_numpy(arr.view(np.int16)).to(device).view(torch.uint16)
    if arr.dtype not in (np.uint8, np.float32):
        arr = arr.astype(np.float32)
    return torch.from_numpy(arr).to(device)

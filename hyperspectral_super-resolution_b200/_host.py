"""Host <-> device marshalling shared by the reference-compatible wrappers.

The reference's call surface is numpy-in / numpy-out; these helpers move such arrays to the
current CUDA device, and back, so the arithmetic always runs in the CUDA kernels.  torch CUDA
tensors pass through untouched (and results stay on the device).
"""
from __future__ import annotations

import numpy as np
import torch


def cuda_device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("hsr_b200 needs a CUDA device (B200, sm_100a): there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def is_numpy_like(a) -> bool:
    return not isinstance(a, torch.Tensor)


def to_device(a, dtype: torch.dtype, device=None) -> torch.Tensor:
    """numpy array / sequence / tensor -> CUDA tensor of ``dtype`` (bool arrays become uint8 0/1)."""
    if isinstance(a, torch.Tensor):
        t = a
        if not t.is_cuda:
            t = t.to(device or cuda_device(), non_blocking=True)
    else:
        arr = np.asarray(a)
        if arr.dtype == np.bool_:
            arr = arr.view(np.uint8)
        if not arr.flags.c_contiguous:
            arr = np.ascontiguousarray(arr)
        if not arr.flags.writeable:
            arr = arr.copy()
        t = torch.from_numpy(arr).to(device or cuda_device(), non_blocking=True)
    if t.dtype == torch.bool and dtype == torch.uint8:
        t = t.view(torch.uint8)
    elif t.dtype != dtype:
        t = t.to(dtype)
    return t


def to_host(t: torch.Tensor, np_dtype=None) -> np.ndarray:
    a = t.detach().cpu().numpy()
    if np_dtype is not None and a.dtype != np_dtype:
        a = a.astype(np_dtype)
    return a

"""Input formats of the synthesis path — call surface of the reference's ``s2_emit/emit_io.py``
(``load_emit_envi_rfl`` :7-16, ``load_emit_wavelengths_from_nc`` :18-33).

The reference reads the ENVI pair that ``nc_to_envi`` wrote with ``spectral.io.envi`` and the wavelength table with h5py;
neither is installed here.  The ENVI reader below is plain numpy (header parser + one ``np.fromfile``; bil / bip / bsq, any
ENVI data type, both byte orders) and returns what ``img.load()`` returns: a ``(lines, samples, bands)`` array.  The
wavelength loader takes whichever HDF5 reader exists (h5py, netCDF4, h5netcdf), imported lazily.  No arithmetic here:
format adapters in front of ``pseudo_s2_srf_integral``.
"""
from __future__ import annotations

import re
from pathlib import Path
from typing import Dict, Optional, Tuple

import numpy as np

_ENVI_DTYPES = {1: "u1", 2: "i2", 3: "i4", 4: "f4", 5: "f8", 12: "u2", 13: "u4", 14: "i8", 15: "u8"}


def read_envi_header(hdr_path) -> Dict[str, str]:
    """``key = value`` pairs of an ENVI header; brace lists (possibly spanning lines) are returned without the braces."""
    text = Path(hdr_path).read_text(errors="replace")
    if not text.lstrip().startswith("ENVI"):
        raise ValueError(f"{hdr_path}: not an ENVI header (missing the 'ENVI' magic line)")
    out: Dict[str, str] = {}
    for m in re.finditer(r"^\s*([^=\n{}]+?)\s*=\s*(\{.*?\}|[^\n]*)", text, flags=re.S | re.M):
        key, val = m.group(1).strip().lower(), m.group(2).strip()
        if val.startswith("{"):
            val = " ".join(val[1:-1].split())
        out[key] = val
    return out


def envi_list(value: str, dtype=float) -> np.ndarray:
    """A brace list of a header (``wavelength``, ``fwhm``, ``map info`` numbers ...) as an array."""
    return np.array([dtype(v) for v in value.split(",") if v.strip() != ""])


def load_emit_envi_rfl(hdr_path: str, bin_path: str, as_float32: bool = True) -> np.ndarray:
    """Load an EMIT reflectance ENVI pair into memory: R (H, W, B)  (reference :7-16)."""
    h = read_envi_header(hdr_path)
    try:
        W, H, B = int(h["samples"]), int(h["lines"]), int(h["bands"])
        code = int(h.get("data type", 4))
    except KeyError as e:
        raise ValueError(f"{hdr_path}: ENVI header lacks {e}") from e
    if code not in _ENVI_DTYPES:
        raise ValueError(f"{hdr_path}: unsupported ENVI data type {code}")
    dt = np.dtype((">" if int(h.get("byte order", 0)) == 1 else "<") + _ENVI_DTYPES[code])
    off = int(h.get("header offset", 0))
    inter = h.get("interleave", "bsq").lower()
    a = np.fromfile(bin_path, dtype=dt, count=H * W * B, offset=off)
    if a.size != H * W * B:
        raise ValueError(f"{bin_path}: {a.size} samples on disk, header says {H} x {W} x {B}")
    if inter == "bil":
        R = a.reshape(H, B, W).transpose(0, 2, 1)
    elif inter == "bip":
        R = a.reshape(H, W, B)
    elif inter == "bsq":
        R = a.reshape(B, H, W).transpose(1, 2, 0)
    else:
        raise ValueError(f"{hdr_path}: unknown interleave {inter!r}")
    R = np.ascontiguousarray(R)
    if as_float32:
        R = R.astype(np.float32, copy=False)
    elif not R.dtype.isnative:
        R = R.astype(R.dtype.newbyteorder("="))
    return R


def _read_h5_pair(nc_path: str, wavelengths_key: str, good_key: str):
    """(wavelengths, good | None) with whichever reader is installed."""
    try:
        import h5py
        with h5py.File(nc_path, "r") as f:  # pragma: no cover  (no HDF5 reader in the build image)
            return f[wavelengths_key][:], (f[good_key][:] if good_key in f else None)
    except ImportError:
        pass

    def by_groups(root, key):  # pragma: no cover
        node = root
        *groups, var = key.split("/")
        for g in groups:
            node = node.groups[g]
        return node.variables[var][:] if var in node.variables else None
    try:
        import netCDF4
        with netCDF4.Dataset(nc_path, "r") as ds:  # pragma: no cover
            return by_groups(ds, wavelengths_key), by_groups(ds, good_key)
    except ImportError:
        pass
    try:
        import h5netcdf
    except ImportError as e:
        raise ImportError("reading the EMIT wavelength table needs h5py, netCDF4 or h5netcdf (none is installed)") from e
    with h5netcdf.File(nc_path, "r") as ds:  # pragma: no cover
        return by_groups(ds, wavelengths_key), by_groups(ds, good_key)


def load_emit_wavelengths_from_nc(nc_path: str, wavelengths_key: str = "sensor_band_parameters/wavelengths",
                                  good_key: str = "sensor_band_parameters/good_wavelengths",
                                  ) -> Tuple[np.ndarray, Optional[np.ndarray]]:
    """(emit_wavelengths_nm float32, good_mask bool | None)  (reference :18-33)."""
    w, good = _read_h5_pair(str(nc_path), wavelengths_key, good_key)
    if w is None:
        raise KeyError(wavelengths_key)
    return np.asarray(w).astype(np.float32), (None if good is None else np.asarray(good).astype(bool))

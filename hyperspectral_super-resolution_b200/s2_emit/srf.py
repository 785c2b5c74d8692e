"""Sentinel-2 spectral response functions: table loading and host-side weight folding.

Call surface of the reference's ``s2_emit/srf.py`` (``DEFAULT_SRF_XLSX_URL`` :6-9, ``S2_BANDS_13``
:11, ``pick_sheet_name`` :13-18, ``load_s2_srf_from_xlsx`` :20-52) plus the new host-side
preparation step of the fused kernel, :func:`srf_fold_weights`.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

DEFAULT_SRF_XLSX_URL = (
    "https://sentiwiki.copernicus.eu/__attachments/1692737/"
    "COPE-GSEG-EOPG-TN-15-0007%20-%20Sentinel-2%20Spectral%20Response%20Functions%202022%20-%203.2.xlsx"
)

S2_BANDS_13 = ["B1", "B2", "B3", "B4", "B5", "B6", "B7", "B8", "B8A", "B9", "B10", "B11", "B12"]

SrfDict = Dict[str, Tuple[np.ndarray, np.ndarray]]


def pick_sheet_name(xl, platform: str = "S2A") -> str:
    """First sheet whose name holds 'Spectral Responses' and the platform (reference srf.py:13-18)."""
    platform = platform.upper()
    for name in xl.sheet_names:
        if "Spectral Responses" in name and platform in name:
            return name
    raise ValueError(f"No sheet containing 'Spectral Responses' and '{platform}' found. Sheets: {xl.sheet_names}")


def load_s2_srf_from_xlsx(xlsx_url: str = DEFAULT_SRF_XLSX_URL, platform: str = "S2A",
                          bands: Optional[List[str]] = None, wavelength_col: str = "SR_WL",
                          col_prefix: Optional[str] = None) -> SrfDict:
    """``{band: (lambda_nm, response)}`` keeping rows with finite, strictly positive response
    (reference srf.py:20-52).  Needs pandas + openpyxl and, for the default URL, network access."""
    import pandas as pd

    bands = bands or S2_BANDS_13
    platform = platform.upper()
    prefix = col_prefix if col_prefix is not None else f"{platform}_SR_AV_"
    xl = pd.ExcelFile(xlsx_url)
    sheet = pick_sheet_name(xl, platform=platform)
    table = xl.parse(sheet)
    lam_all = pd.to_numeric(table[wavelength_col], errors="coerce").to_numpy()
    srf: SrfDict = {}
    for band in bands:
        column = f"{prefix}{band}"
        if column not in table.columns:
            raise KeyError(f"Column '{column}' not found in sheet '{sheet}'.")
        resp = pd.to_numeric(table[column], errors="coerce").to_numpy()
        keep = np.isfinite(lam_all) & np.isfinite(resp) & (resp > 0)
        srf[band] = (lam_all[keep].astype(float), resp[keep].astype(float))
    return srf


# centre / width (nm) of the S2 bands: Pairs_EMIT_S2_demo-2.ipynb cell 57 (+ B10, the cirrus band)
_S2_CENTRE_WIDTH = {
    "B1": (443, 20), "B2": (490, 65), "B3": (560, 35), "B4": (665, 30), "B5": (705, 15), "B6": (740, 15),
    "B7": (783, 20), "B8": (842, 115), "B8A": (865, 20), "B9": (945, 20), "B10": (1375, 30), "B11": (1610, 90),
    "B12": (2190, 180),
}


def synthetic_s2_srf(bands: Optional[Sequence[str]] = None) -> SrfDict:
    """Synthetic SRF tables in the format of :func:`load_s2_srf_from_xlsx` (no network in the build image).

    Super-Gaussian ``exp(-0.5 ((lam - c) / (w / 2.355))**4)`` on a 1 nm grid 300..2600 nm, zeroed below
    1e-3 and filtered ``resp > 0`` exactly like srf.py:47-50 filters the real table.
    """
    lam = np.arange(300.0, 2601.0, 1.0)
    out: SrfDict = {}
    for band in (bands or S2_BANDS_13):
        c, w = _S2_CENTRE_WIDTH[band]
        resp = np.exp(-0.5 * ((lam - c) / (w / 2.355)) ** 4)
        resp[resp < 1e-3] = 0.0
        keep = np.isfinite(resp) & (resp > 0)
        out[band] = (lam[keep].astype(float), resp[keep].astype(float))
    return out


def trapezoid_weights(x: np.ndarray) -> np.ndarray:
    """Node weights of np.trapz on the grid x: sum_i d_i (y_i + y_{i+1}) / 2 == sum_b tw[b] y[b]."""
    x = np.asarray(x, dtype=np.float64)
    tw = np.zeros_like(x)
    if x.size >= 2:
        d = np.diff(x)
        tw[:-1] += 0.5 * d
        tw[1:] += 0.5 * d
    return tw


def srf_fold_weights(emit_w, srf_dict: SrfDict, good_mask=None, *, fill: float = -9999.0):
    """Fold interpolation, good-band mask, trapezoid rule and normalisation of
    ``pseudo_s2_srf_integral`` (reference s2_emit/synth.py:25,33-35,41-43) into one matrix.

    Returns ``(W, names, none_bands, fill_out)``:
      W          float32 [B, K] with ``band_k(pixel) = sum_b R[pixel, b] * W[b, k]``
      names      the K bands that are not ``None`` in the reference (dict order kept)
      none_bands bands whose response is all-zero on the EMIT grid (synth.py:37-39)
      fill_out   float32 [K] = fill * sum_b W[b, k] in float64 (the value an all-fill pixel integrates to)
    """
    lam = np.asarray(emit_w).astype(float)            # synth.py:25
    if lam.ndim != 1:
        raise ValueError(f"emit_w must be (B,). Got {lam.shape}")
    good = None if good_mask is None else np.asarray(good_mask).astype(float)
    tw = trapezoid_weights(lam)
    cols, names, none_bands = [], [], []
    for band, (lam_srf, rsp_srf) in srf_dict.items():
        rsp = np.interp(lam, lam_srf, rsp_srf, left=0.0, right=0.0)   # synth.py:33
        if good is not None:
            rsp = rsp * good                                          # synth.py:34-35
        if np.all(rsp == 0):                                          # synth.py:37-39
            none_bands.append(band)
            continue
        den = float(np.sum(tw * rsp))                                 # np.trapz(rsp, x=lam), synth.py:42
        cols.append(rsp * tw / (den + 1e-32))                         # synth.py:41,43
        names.append(band)
    if cols:
        W64 = np.stack(cols, axis=1)
    else:
        W64 = np.zeros((lam.shape[0], 0))
    fill_out = (W64.sum(axis=0) * float(fill)).astype(np.float32)
    return np.ascontiguousarray(W64.astype(np.float32)), names, none_bands, fill_out

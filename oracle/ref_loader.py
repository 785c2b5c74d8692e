"""Load the reference's OWN arithmetic functions from /root/reference (test infrastructure).

Used by ``tests/golden/make_golden.py`` (to produce the committed golden vectors) and by CPU tests
that cross-check the oracle against the live reference when it is mounted.  /root/reference does
not exist on the GPU box: nothing in the ``-m gpu`` tests, ``smoke()`` or ``bench.py`` calls this.

Package imports of the reference fail in this image (h5py / rasterio / POT / xarray / hytools are
absent, and ``s2_emit/poly_regression.py`` runs a script at import time), so only the
``FunctionDef`` nodes of the pure-numpy functions are compiled, with numpy as their only global.
No reference source is copied into the repo.
"""
from __future__ import annotations

import ast
import os
import warnings
from types import SimpleNamespace

import numpy as np

REFERENCE_ROOT = os.environ.get("HSR_REFERENCE_ROOT", "/root/reference")

_WANTED = {
    "EMIT_data/emit_tools.py": ["apply_glt"],
    "s2_emit/synth.py": ["pseudo_s2_srf_integral", "pseudo_s2_rgb"],
    "s2_emit/poly_regression.py": ["fit_ot_poly_rgb", "apply_poly_rgb"],
    "s2_emit/color.py": ["apply_shared_percentile_stretch", "robust_norm", "robust_norm_rgb", "ot_match_rgb_sinkhorn_pot",
                        "_hist_match_channel", "histogram_match_rgb"],
    "tiles_helpers/utils.py": ["is_black_mask", "_subsample_bands_evenly"],
}


def available() -> bool:
    return all(os.path.exists(os.path.join(REFERENCE_ROOT, f)) for f in _WANTED)


def _extract(path: str, names, ot_module=None):
    with open(path) as fh:
        tree = ast.parse(fh.read(), filename=path)
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    missing = set(names) - {n.name for n in body}
    if missing:
        raise RuntimeError(f"{path}: functions not found: {sorted(missing)}")
    module = ast.Module(body=body, type_ignores=[])
    from typing import Dict, List, Optional, Tuple
    if ot_module is None:
        from . import ot as ot_module    # POT is absent: its published dist / sinkhorn, restated (parity unpinned)
    env = {"np": np, "Dict": Dict, "Tuple": Tuple, "Optional": Optional, "List": List, "ot": ot_module}
    exec(compile(module, path, "exec"), env)
    return {n: env[n] for n in names}


def load(ot_module=None) -> SimpleNamespace:
    """Namespace with the reference's apply_glt, pseudo_s2_srf_integral, pseudo_s2_rgb,
    fit_ot_poly_rgb (its ``ot.dist`` / ``ot.sinkhorn`` calls resolve to oracle/ot.py: POT is absent — or to
    ``ot_module``, the real POT, when tests/golden/make_golden_pot.py runs where it is installed),
    apply_poly_rgb, apply_shared_percentile_stretch."""
    if not available():
        raise FileNotFoundError(f"reference not mounted at {REFERENCE_ROOT}")
    fns = {}
    for rel, names in _WANTED.items():
        fns.update(_extract(os.path.join(REFERENCE_ROOT, rel), names, ot_module))
    warnings.filterwarnings("ignore", message=".*trapz.*", category=DeprecationWarning)
    return SimpleNamespace(**fns)

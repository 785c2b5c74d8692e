"""oracle/ — CPU restatement of the reference's arithmetic for the pair-synthesis hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as the
checker (or as the timed CPU baseline) — never on the product path, which must fail loudly when
the CUDA library is missing.

The reference (martasumyk/hyperspectral_super-resolution) is pure Python/numpy, so the
restatement is numpy too (float64 exactly where the reference is float64).  Each function cites
the reference file:line it follows.

Parity pinning: the reference ships no tests, fixtures or golden vectors.  The
oracle is therefore pinned against OUTPUTS OF THE REFERENCE ITSELF: ``tests/golden/make_golden.py``
imports the reference's own functions from /root/reference (``apply_glt``,
``pseudo_s2_srf_integral``, ``pseudo_s2_rgb``, ``apply_poly_rgb``, ``fit_ot_poly_rgb``'s
identity branch) on seeded inputs and commits inputs + outputs as ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every oracle function against them (and against the live
reference when /root/reference is mounted).  Two pieces have no callable reference and are
pinned indirectly: the in-bounds rule of ``nc_to_envi`` (inline in a 700-line I/O function;
cross-checked against ``apply_glt`` on in-range GLTs, where they must agree bit for bit) and
``np.polyfit`` (numpy's own, the function the reference calls).  The Sinkhorn/OT target stage calls POT,
which is absent and unpinned: ``oracle/ot.py`` restates POT's published ``dist`` / ``sinkhorn_knopp`` and is
injected as ``ot`` into the reference's own ``fit_ot_poly_rgb`` — PARITY UNPINNED for those two functions
(see that module's header), pinned for everything around them.  ``oracle/color.py`` (percentile stretch) is
pinned against the reference's ``apply_shared_percentile_stretch``.
"""
from . import color, glt, ot, poly, resample, srf, tiles, warp  # noqa: F401

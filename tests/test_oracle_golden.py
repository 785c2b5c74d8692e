"""The oracle (oracle/*.py) against the golden vectors produced by the reference's own functions
(tests/golden/make_golden.py) and, when /root/reference is mounted, against the live reference."""
import warnings

import numpy as np
import pytest

from oracle import color as ocolor
from oracle import glt as oglt
from oracle import poly as opoly
from oracle import ref_loader
from oracle import srf as osrf

warnings.filterwarnings("ignore", category=RuntimeWarning)
warnings.filterwarnings("ignore", category=DeprecationWarning)


def _srf_from_golden(g):
    names = [str(n) for n in g["names"]]
    return names, {b: (g[f"lam_{b}"], g[f"rsp_{b}"]) for b in names}


def test_glt_oracle_matches_reference_apply_glt(golden):
    g = golden("glt_apply_glt.npz")
    raw, glt = g["raw"], g["glt"]
    # xarray flavour restatement: bit-exact incl. NaN payloads
    out = oglt.apply_glt(raw, glt)
    assert out.dtype == np.float32
    assert np.array_equal(out.view(np.int32), g["ortho"].view(np.int32))
    # production (nc_to_envi) restatement agrees bit for bit on an in-range GLT
    out2, valid, diag = oglt.glt_ortho(raw, glt[..., 0], glt[..., 1])
    assert np.array_equal(out2.view(np.int32), g["ortho"].view(np.int32))
    assert diag["valid_glt_dropped_oob"] == 0
    assert diag["valid_glt_count"] == int(valid.sum())
    assert np.array_equal(valid, np.all(glt != 0, axis=-1))
    # 2-D plane flavour
    pl = oglt.glt_plane(raw[..., 40], glt[..., 0], glt[..., 1])
    assert np.array_equal(pl.view(np.int32), g["plane"][..., 0].view(np.int32))


def test_glt_oracle_inbounds_rule():
    # emit_proj.py:691-703: zeros, negatives, too-large and NaN entries are all dropped
    raw = np.arange(3 * 4 * 2, dtype=np.float32).reshape(3, 4, 2)
    gx = np.array([[1, 4, 5, 0], [2, -1, 3, np.nan]], dtype=np.float64)
    gy = np.array([[1, 3, 1, 2], [0, 2, 4, 1]], dtype=np.float64)
    out, valid, diag = oglt.glt_ortho(raw, gx, gy)
    assert valid.tolist() == [[True, True, False, False], [False, False, False, False]]
    assert np.array_equal(out[0, 0], raw[0, 0]) and np.array_equal(out[0, 1], raw[2, 3])
    assert np.all(out[~valid] == np.float32(-9999.0))
    assert diag == {"raw_shape_yx": [3, 4], "valid_glt_count": 5, "valid_glt_inbounds_count": 2,
                    "valid_glt_dropped_oob": 3}


def test_glt_oracle_transposed_raw():
    rng = np.random.default_rng(3)
    raw = rng.random((5, 7, 6), dtype=np.float32)            # logical (y, x, b)
    phys = np.ascontiguousarray(raw.transpose(1, 0, 2))     # file order (crosstrack, downtrack, b)
    gx = rng.integers(0, 8, size=(9, 4)).astype(np.int32)
    gy = rng.integers(0, 6, size=(9, 4)).astype(np.int32)
    a, va, _ = oglt.glt_ortho(raw, gx, gy)
    b, vb, _ = oglt.glt_ortho(phys, gx, gy, transpose_raw_yx=True)
    assert np.array_equal(a, b) and np.array_equal(va, vb)


def test_srf_oracle_matches_reference(golden):
    g = golden("srf_pseudo_s2.npz")
    names, srf = _srf_from_golden(g)
    for tag, good in (("good", g["good"]), ("all", None)):
        out = osrf.pseudo_s2_srf_integral(g["cube"], g["emit_w"], srf, good, rows_per_slab=3)
        assert list(out.keys()) == names
        for b in names:
            if bool(g[f"none_{tag}_{b}"]):
                assert out[b] is None
            else:
                ref = g[f"out_{tag}_{b}"]
                assert out[b].dtype == np.float64
                np.testing.assert_allclose(out[b], ref, rtol=1e-13, atol=0, equal_nan=True)
    out = osrf.pseudo_s2_srf_integral(g["cube"], g["emit_w"], srf, g["good"])
    assert out["B10"] is None                                   # cirrus band sits in a masked window
    assert np.isnan(out["B2"][1, 2]) and np.isnan(out["B12"][1, 2])   # NaN in a zero-weight band poisons all
    assert np.isnan(out["B4"][3, 4])
    assert out["B4"][4, 1] == np.inf and np.isnan(out["B3"][4, 1])   # Inf under a non-zero weight stays Inf
    assert out["B11"][5, 2] == -np.inf and np.isnan(out["B12"][5, 2])
    assert abs(out["B3"][0, 0] + 9999.0) < 1e-6                 # fill integrates to ~ -9999
    np.testing.assert_allclose(osrf.pseudo_s2_rgb(out), g["rgb"], rtol=1e-13, equal_nan=True)
    with pytest.raises(ValueError):
        osrf.pseudo_s2_rgb(out, order=("B10", "B3", "B2"))
    with pytest.raises(ValueError):
        osrf.pseudo_s2_srf_integral(g["cube"][0], g["emit_w"], srf)
    with pytest.raises(ValueError):
        osrf.pseudo_s2_srf_integral(g["cube"], g["emit_w"][:-1], srf)


def test_poly_oracle_matches_reference(golden):
    g = golden("poly_apply_fit.npz")
    img, mask = g["img"], g["mask"]
    for key, coeffs, m in (("out2_mask", g["coeffs2"], mask), ("out2_nomask", g["coeffs2"], None),
                           ("out4_mask", g["coeffs4"], mask)):
        out = opoly.apply_poly_rgb(img, coeffs, m)
        assert out.dtype == np.float32
        assert np.array_equal(out, g[key], equal_nan=True), key
    # unmasked pixels are still clipped (poly_regression.py:84)
    assert g["out2_mask"][~mask].max() <= 1.0 and g["out2_mask"][~mask & np.isfinite(img).all(-1)].min() >= 0.0
    # paired fit == per-channel np.polyfit on the kept pixels
    for deg, key in ((2, "fit2"), (4, "fit4")):
        c = opoly.fit_poly_rgb_paired(img, g["yimg"], mask, deg)
        np.testing.assert_allclose(c, g[key], rtol=1e-12)
        planes_x = np.moveaxis(img, -1, 0)
        planes_y = np.moveaxis(g["yimg"], -1, 0)
        keep = mask & np.isfinite(img).all(-1) & np.isfinite(g["yimg"]).all(-1)
        c2 = opoly.polyfit_paired(planes_x, planes_y, keep, deg)
        np.testing.assert_allclose(c2, g[key], rtol=1e-12)
    # < 200 samples -> identity (poly_regression.py:38-41)
    ident = opoly.fit_poly_rgb_paired(img, g["yimg"], g["small_mask"], 2)
    assert np.array_equal(ident, g["ident"]) and np.array_equal(ident, np.array([[0, 1, 0]] * 3, float))
    # planar apply agrees with the interleaved one
    pl = opoly.apply_poly_planes(np.moveaxis(img, -1, 0), g["coeffs2"], mask)
    assert np.array_equal(np.moveaxis(pl, 0, -1), g["out2_mask"], equal_nan=True)


def test_color_oracle_matches_reference(golden):
    """oracle/color.py against the golden outputs of the reference's apply_shared_percentile_stretch
    (s2_emit/color.py:25-34), and the written-out percentile against numpy's on awkward sample sets."""
    from oracle import color as ocolor

    g = golden("color_stretch.npz")
    with np.errstate(invalid="ignore"):
        assert np.array_equal(ocolor.apply_shared_percentile_stretch(g["img"], g["mask"]), g["out"])
        assert np.array_equal(ocolor.apply_shared_percentile_stretch(g["img"], g["mask"], 1, 99.5), g["out_1_995"])
        assert np.array_equal(ocolor.apply_shared_percentile_stretch(g["img"], g["tiny_mask"]), g["tiny"])
        nout = ocolor.apply_shared_percentile_stretch(g["nanimg"], g["nmask"])
    assert np.array_equal(nout, g["nout"], equal_nan=True) and np.isnan(nout[..., 1]).all()
    assert g["out"].dtype == np.float32 and g["out"].min() == 0.0 and g["out"].max() == 1.0
    assert np.array_equal(ocolor.shared_percentile_limits(g["img"], g["mask"]), g["limits"])
    rng = np.random.default_rng(3)
    for n in (1, 2, 3, 7, 50, 51, 1000, 4097):
        for kind in range(4):
            v = rng.normal(0, 1, n).astype(np.float32)
            if kind == 1:
                v = np.round(v * 2) / 2                                      # ties
            if kind == 2 and n > 2:
                v[rng.integers(n)] = np.inf
                v[rng.integers(n)] = -np.inf
            if kind == 3 and n > 3:
                v[rng.integers(n)] = np.nan
            for q in ([2, 98], [0, 100], [50, 99.9], [37.5, 62.5]):
                with np.errstate(invalid="ignore"):
                    want = np.percentile(v, q)
                    got = ocolor.percentile_linear_sorted(v, q)
                assert np.array_equal(got, want, equal_nan=True), (n, kind, q)
    with pytest.raises(IndexError):
        ocolor.apply_shared_percentile_stretch(g["img"], np.zeros_like(g["mask"]))


def test_tiles_oracle_and_host_logic_match_reference(golden):
    """oracle/tiles.py and the host-side band picker against the golden outputs of the reference's is_black_mask /
    _subsample_bands_evenly (tiles_helpers/utils.py:201-220, :444-458)."""
    from hsr_b200.tiles_helpers import _subsample_bands_evenly
    from oracle import tiles as otiles

    g = golden("tiles.npz")
    tile = g["tile"]
    assert np.array_equal(otiles.is_black_mask(tile, nodata=-9999.0), g["black_nd"])
    assert np.array_equal(otiles.is_black_mask(tile), g["black_none"])
    assert np.array_equal(otiles.is_black_mask(tile, nodata=0.5, masked_val=0.25, nodata_atol=0.3, zero_atol=0.05),
                          g["black_custom"])
    assert g["black_nd"][0, 0] and not g["black_nd"][0, 1] and g["black_nd"][1, 1] and not g["black_nd"][1, 2]
    assert g["black_nd"][3, 0] and not g["black_nd"][3, 1] and not g["black_nd"][4, 0]
    for key in [k for k in g.files if k.startswith("idx_")]:
        _, n, k = key.split("_")
        assert np.array_equal(otiles.subsample_bands_evenly(int(n), int(k)), g[key]), key
        assert np.array_equal(_subsample_bands_evenly(int(n), int(k)), g[key]), key
    # quantisation (tiles_helpers/utils.py:357-371): known answers
    x = np.array([0.0, 0.12345, 1.0, 6.5534, 6.5535, 7.0, -0.01, -9999.0, np.nan, np.inf, 0.00005, 0.00015, 3e5, -3e5],
                 dtype=np.float32)
    q = otiles.quantize_emit_u16(x, nodata=-9999.0)
    assert q.dtype == np.uint16
    # rint ties to even in float32; 3e5 * 1e4 overflows int32 -> INT_MIN on x86 -> clipped to 0; fill / NaN / Inf -> 65535
    assert q.tolist() == [0, 1234, 10000, 65534, 65534, 65534, 0, 65535, 65535, 65535, 0, 2, 0, 0]


def test_ot_oracle_matches_reference_function(golden):
    """oracle/ot.fit_ot_poly_rgb == the reference's own fit_ot_poly_rgb (golden; POT's dist / sinkhorn restated
    in oracle/ot.py — parity unpinned for those two), and the Sinkhorn restatement has the properties POT's has."""
    from oracle import ot as oot

    g = golden("ot_fit.npz")
    for key in [k for k in g.files if k.startswith("coeffs_")]:
        _, d, n, sd = key.split("_")
        c = oot.fit_ot_poly_rgb(g["src"], g["ref"], g["mask"], deg=int(d[1:]), n_samples=int(n[1:]), seed=int(sd[1:]))
        assert np.array_equal(c, g[key]), key
    assert np.array_equal(oot.fit_ot_poly_rgb(g["src"], g["ref"], g["small"], deg=3), g["ident"])
    assert np.array_equal(g["ident"], np.array([[0, 0, 1, 0]] * 3, float))
    X, Y = g["X"], g["Y"]
    assert np.array_equal(oot.barycentric_targets(X, Y), g["ybar"])
    M = oot.dist(X, Y)
    direct = ((X[:, None, :] - Y[None, :, :]) ** 2).sum(-1)
    assert M.shape == (500, 430) and M.min() >= 0 and np.allclose(M, direct, atol=1e-14)
    a, b = np.full(500, 1 / 500), np.full(430, 1 / 430)
    P, info = oot.sinkhorn_knopp(a, b, M, 0.05, 300, 1e-6, log=True)
    assert info["err"][-1] < 1e-6 and info["niter"] % 10 == 0 and not info["numerical"]
    assert np.allclose(P.sum(0), b, atol=2e-6) and np.allclose(P.sum(1), a, atol=1e-12)   # u was updated last
    # barycentric targets are convex combinations of the reference samples
    assert (g["ybar"] >= Y.min(0) - 1e-12).all() and (g["ybar"] <= Y.max(0) + 1e-12).all()
    with np.errstate(all="ignore"):
        P2, info2 = oot.sinkhorn_knopp(a, b, M * 1e6, 0.05, 50, 1e-9, log=True)    # K underflows to 0: roll back
    assert info2["numerical"] and info2["niter"] == 0 and np.isfinite(P2).all()


def test_ot_colour_transfer_oracle_matches_reference_function(golden):
    """oracle/color.ot_match_rgb_sinkhorn_pot == the reference's own function (s2_emit/color.py:63-116; golden made by
    tests/golden/make_golden_color.py with oracle/ot.py injected as POT — parity unpinned for dist / sinkhorn only)."""
    g = golden("color_ot_match.npz")
    with np.errstate(invalid="ignore"):
        assert np.array_equal(ocolor.ot_match_rgb_sinkhorn_pot(g["src"], g["ref"], g["mask"], n_samples=400, seed=0),
                              g["out_n400_s0"], equal_nan=True)
        assert np.array_equal(ocolor.ot_match_rgb_sinkhorn_pot(g["src"], g["ref"], g["mask"], n_samples=100000, seed=2),
                              g["out_n100000_s2"], equal_nan=True)
        assert np.array_equal(ocolor.ot_match_rgb_sinkhorn_pot(g["src"], g["ref"], g["mask"], n_samples=300, reg=0.1,
                                                               numItermax=20, stopThr=0.0, seed=5),
                              g["out_reg01_it20"], equal_nan=True)
    assert np.array_equal(ocolor.ot_match_rgb_sinkhorn_pot(g["src"], g["ref"], g["one"]), g["src"], equal_nan=True)
    assert np.array_equal(g["out_one"], g["src"], equal_nan=True)
    out = g["out_n400_s0"]
    m = g["mask"]
    assert out.dtype == np.float32 and np.array_equal(out[~m], g["src"][~m], equal_nan=True)      # outside the mask: copied
    assert np.isnan(out[3, 4]).all()                                   # a NaN channel poisons the whole pixel (x @ A)
    fin = np.isfinite(out[m]).all(axis=1)
    assert (out[m][fin] >= 0).all() and (out[m][fin] <= 1).all()


def test_robust_norm_oracle_matches_reference_functions(golden):
    """oracle/color.robust_norm / robust_norm_rgb == the reference's own functions (s2_emit/color.py:6-23)."""
    g = golden("color_robust.npz")
    with np.errstate(invalid="ignore"):
        assert np.array_equal(ocolor.robust_norm(g["xn"]), g["rn"], equal_nan=True)
        assert np.array_equal(ocolor.robust_norm(g["xn"], 5, 90), g["rn_5_90"], equal_nan=True)
        assert np.array_equal(ocolor.robust_norm(g["img"]), g["rn_cube"], equal_nan=True)
        assert np.array_equal(ocolor.robust_norm_rgb(g["img"], g["mask"]), g["rgb"], equal_nan=True)
        assert np.array_equal(ocolor.robust_norm_rgb(g["img"], g["mask"], 1, 99), g["rgb_1_99"], equal_nan=True)
    assert g["rn"].dtype == np.float64 and g["rgb"].dtype == np.float64
    assert np.array_equal(np.isnan(g["rn"]), np.isnan(g["xn"])) and g["rn"][0, 0] == 1.0      # NaN kept, +Inf clips to 1
    assert np.isnan(g["rgb"][~g["mask"]]).all() and not np.isnan(g["rgb"][g["mask"]]).any()


def test_histogram_match_oracle_matches_reference_function(golden):
    """oracle/color.histogram_match_rgb == the reference's own function (s2_emit/color.py:36-63)."""
    g = golden("color_histmatch.npz")
    assert np.array_equal(ocolor.histogram_match_rgb(g["src"], g["ref"], g["mask"]), g["out"])
    assert np.array_equal(ocolor.histogram_match_rgb(g["src"], g["ref"], g["few"]), g["out_few"])
    out, m = g["out"], g["mask"]
    assert out.dtype == np.float32 and out.min() >= 0 and out.max() <= 1
    assert np.array_equal(out[~m], np.clip(g["src"][~m], 0, 1))              # outside the mask: only clipped
    for c in range(3):                                                       # matched values are reference values or between
        assert out[..., c][m].min() >= g["ref"][..., c][m].min() and out[..., c][m].max() <= g["ref"][..., c][m].max()


@pytest.mark.skipif(not ref_loader.available(), reason="reference not mounted (GPU box)")
def test_oracle_against_live_reference():
    from hsr_b200 import synthetic
    from hsr_b200.s2_emit.srf import synthetic_s2_srf

    ref = ref_loader.load()
    rng = np.random.default_rng(99)
    w = synthetic.emit_wavelengths()
    good = synthetic.good_band_mask(w)
    for seed in range(3):
        Hr, Wr = 9 + seed, 13 - seed
        raw = synthetic.raw_cube_bits_np((Hr, Wr, 285), seed=seed, good=good)
        gx, gy = synthetic.rotation_glt(Hr, Wr, 10.0 + 20 * seed)
        gx[rng.random(gx.shape) < 0.03] = 0
        glt = np.stack([gx, gy], -1).astype(int)
        a = ref.apply_glt(raw, glt)
        b, _, _ = oglt.glt_ortho(raw, gx, gy)
        assert np.array_equal(a.view(np.int32), b.view(np.int32))
        srf = synthetic_s2_srf()
        ps_ref = ref.pseudo_s2_srf_integral(a, w, srf, good)
        ps_or = osrf.pseudo_s2_srf_integral(a, w, srf, good, rows_per_slab=4)
        for band in srf:
            if ps_ref[band] is None:
                assert ps_or[band] is None
            else:
                np.testing.assert_allclose(ps_or[band], ps_ref[band], rtol=1e-13, equal_nan=True)
        rgb = rng.uniform(-0.1, 1.2, size=(12, 10, 3)).astype(np.float32)
        coeffs = rng.normal(0, 0.6, size=(3, 3 + seed))
        mask = rng.random((12, 10)) < 0.5
        assert np.array_equal(ref.apply_poly_rgb(rgb, coeffs, mask), opoly.apply_poly_rgb(rgb, coeffs, mask))
        assert np.array_equal(ref.apply_poly_rgb(rgb, coeffs), opoly.apply_poly_rgb(rgb, coeffs))
        from oracle import color as ocolor
        from oracle import ot as oot
        big = rng.random((30, 25, 3)).astype(np.float32)
        bm = rng.random((30, 25)) < 0.7
        assert np.array_equal(ref.fit_ot_poly_rgb(big, big ** 2, bm, deg=2, n_samples=300, seed=seed),
                              oot.fit_ot_poly_rgb(big, big ** 2, bm, deg=2, n_samples=300, seed=seed))
        assert np.array_equal(ref.apply_shared_percentile_stretch(rgb, mask, 2 + seed, 98 - seed),
                              ocolor.apply_shared_percentile_stretch(rgb, mask, 2 + seed, 98 - seed))


def test_sinkhorn_fixed_point_against_independent_logdomain_solve():
    """until POT itself can pin oracle/ot.py, check its restated ot.dist / sinkhorn_knopp against an
    independent formulation.  The entropic OT optimum is unique, so ANY correct Sinkhorn converges to the same plan:
    oracle (scaling form, float64, run to 1e-14) vs log-domain long double, 64 x 64 and a ragged 48 x 81."""
    from oracle import ot as oot
    from oracle.ot_independent import logdomain_sinkhorn_longdouble as _logdomain_sinkhorn_longdouble

    rng = np.random.default_rng(42)
    for ns, nt in ((64, 64), (48, 81)):
        X = rng.random((ns, 3))
        Y = rng.random((nt, 3)) ** 1.3 * 0.8 + 0.1
        Pref, Mref = _logdomain_sinkhorn_longdouble(X, Y, 0.05, 4000)
        assert abs(float(Pref.sum()) - 1.0) < 1e-15
        assert np.abs(Pref.sum(1) - 1.0 / ns).max() < 1e-17 and np.abs(Pref.sum(0) - 1.0 / nt).max() < 1e-15   # converged
        M = oot.dist(X, Y)
        assert np.abs(M - Mref.astype(np.float64)).max() < 1e-15                                   # ot.dist restatement
        a, b = np.full(ns, 1.0 / ns), np.full(nt, 1.0 / nt)
        P, info = oot.sinkhorn_knopp(a, b, M, 0.05, numItermax=20000, stopThr=1e-15, log=True)
        assert info["err"][-1] < 1e-15 and not info["numerical"]
        assert np.abs(P - Pref.astype(np.float64)).max() < 1e-13 * float(Pref.max()) + 1e-18, (ns, nt)
        # at the reference's own stopping rule (300 iterations / 1e-6 on the column marginal) the plan is within the
        # stopping tolerance of that optimum, its row marginal is exact (u is updated last) and it is non-negative
        P6, info6 = oot.sinkhorn_knopp(a, b, M, 0.05, numItermax=300, stopThr=1e-6, log=True)
        assert info6["err"][-1] < 1e-6 and (P6 >= 0).all()
        assert np.abs(P6.sum(1) - a).max() < 1e-15 and np.linalg.norm(P6.sum(0) - b) < 1e-6
        assert np.abs(P6 - Pref.astype(np.float64)).sum() < 1e-4
        ybar = oot.barycentric_targets(X, Y, 0.05, 20000, 1e-15)
        want = (Pref @ Y.astype(np.longdouble)) / Pref.sum(1, keepdims=True)
        assert np.abs(ybar - want.astype(np.float64)).max() < 1e-12


def test_warp_oracle_point_kernels_known_answers():
    """oracle/warp.py "nearest" / "average" (GWKNearest / GWKAverageOrMode restated; parity with GDAL unpinned) against what
    they must give where the answer is known: the identity warp, a whole-pixel shift, snapped integer-ratio blocks (the
    block mean — the geometry nc_to_envi produces, emit_proj.py:794-797), half-covered footprints, nodata per band."""
    from oracle import resample as oresample
    from oracle import warp as owarp

    rng = np.random.default_rng(11)
    src = rng.random((12, 18, 3)).astype(np.float32)
    sgt = (500000.0, 10.0, 0.0, 4000000.0, 0.0, -10.0)
    for kernel in ("nearest", "average"):
        assert np.array_equal(owarp.warp(src, sgt, sgt, 12, 18, utm=False, nodata=None, kernel=kernel), src)
        shifted = owarp.warp(src, sgt, (500020.0, 10.0, 0.0, 3999990.0, 0.0, -10.0), 11, 16, utm=False, nodata=None, kernel=kernel)
        assert np.array_equal(shifted, src[1:, 2:])                                 # two pixels right, one down
    # 6 x 6 blocks on the snapped grid = block mean; 3 x 2 blocks too
    a = owarp.warp(src, sgt, (500000.0, 60.0, 0.0, 4000000.0, 0.0, -60.0), 2, 3, utm=False, nodata=None, kernel="average")
    assert np.allclose(np.transpose(a, (2, 0, 1)), oresample.downsample_to_grid(np.transpose(src, (2, 0, 1)), 6), rtol=1e-6, atol=1e-7)
    b = owarp.warp(src, sgt, (500000.0, 20.0, 0.0, 4000000.0, 0.0, -30.0), 4, 9, utm=False, nodata=None, kernel="average")
    want = src.reshape(4, 3, 9, 2, 3).astype(np.float64).mean(axis=(1, 3))
    assert np.allclose(b, want, rtol=1e-6, atol=1e-7)
    # a destination pixel of 1.5 source pixels, offset by half a pixel: covers [0.5, 2.0) -> weights 0.5, 1.0 per axis
    c = owarp.warp(src, sgt, (500005.0, 15.0, 0.0, 3999995.0, 0.0, -15.0), 1, 1, utm=False, nodata=None, kernel="average")
    w = np.array([0.5, 1.0])
    want = (src[0:2, 0:2].astype(np.float64) * (w[:, None, None] * w[None, :, None])).sum((0, 1)) / (w.sum() ** 2)
    assert np.allclose(c[0, 0], want, rtol=1e-6)
    # nodata: skipped per band in the average, propagated by nearest; a footprint of nothing but nodata stays nodata
    nd = src.copy()
    nd[0:6, 0:6, 1] = -9999.0
    nd[0, 0, 0] = -9999.0
    d = owarp.warp(nd, sgt, (500000.0, 60.0, 0.0, 4000000.0, 0.0, -60.0), 2, 3, utm=False, nodata=-9999.0, kernel="average")
    assert d[0, 0, 1] == -9999.0 and np.isclose(d[0, 0, 0], (src[0:6, 0:6, 0].astype(np.float64).sum() - src[0, 0, 0]) / 35.0, rtol=1e-6)
    e = owarp.warp(nd, sgt, sgt, 12, 18, utm=False, nodata=-9999.0, kernel="nearest")
    assert e[0, 0, 0] == -9999.0 and e[3, 3, 1] == -9999.0 and e[3, 3, 0] == src[3, 3, 0]
    # outside the source: untouched (dst nodata / 0)
    f = owarp.warp(src, sgt, (499000.0, 10.0, 0.0, 4000000.0, 0.0, -10.0), 2, 2, utm=False, nodata=None, kernel="nearest")
    assert (f == 0).all()

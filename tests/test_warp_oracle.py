"""CPU tests of the general grid warp: the oracle's transverse Mercator against known answers, the
host-side extent arithmetic of ``hsr_b200.EMIT_data.warp`` (the reference's ``_compute_te``, emit_proj.py:354-382)
against the oracle's restatement, and the resampling weights.  No GPU: the kernel itself is checked in
tests/test_gpu_warp.py.  Parity with GDAL / PROJ is unpinned (oracle/warp.py header)."""

import numpy as np
import pytest

from hsr_b200.EMIT_data import warp as hwarp
from oracle import warp as owarp


def test_utm_known_answers():
    # a zone spans +-3 deg of its central meridian: on the equator its edges are at the classic false-easting limits
    assert owarp.utm_forward(0.0, 0.0, 31)[0] == pytest.approx(166021.4431, abs=1e-3)
    assert owarp.utm_forward(6.0, 0.0, 31)[0] == pytest.approx(833978.5569, abs=1e-3)
    # on the central meridian northing = k0 * meridian arc; WGS-84 arc from the equator to 45 N = 4 984 944.378 m
    x, y = owarp.utm_forward(3.0, 45.0, 31)
    assert x == pytest.approx(500000.0, abs=1e-9) and y == pytest.approx(0.9996 * 4984944.378, abs=2e-3)
    # CN Tower, Toronto: 43.642566 N, 79.387139 W -> 17T 630084 E 4833439 N (quoted to the metre)
    x, y = owarp.utm_forward(-79.387139, 43.642566, 17)
    assert abs(x - 630084) < 1.0 and abs(y - 4833439) < 1.0
    # southern hemisphere: false northing 10 000 km
    assert owarp.utm_forward(3.0, -45.0, 31, True)[1] == pytest.approx(10000000.0 - 0.9996 * 4984944.378, abs=2e-3)


def test_utm_round_trip_and_host_matches_oracle():
    rng = np.random.default_rng(0)
    lon = rng.uniform(-180, 180, 500)
    lat = rng.uniform(-80, 84, 500)
    for lo, la in zip(lon, lat):
        zone = min(int((lo + 180) // 6) + 1, 60)
        south = la < 0
        x, y = owarp.utm_forward(lo, la, zone, south)
        lo2, la2 = owarp.utm_inverse(x, y, zone, south)
        assert abs(lo2 - lo) < 1e-11 and abs(la2 - la) < 1e-11
        hx, hy = hwarp.utm_forward(lo, la, zone, south)                 # host numpy (real arithmetic) vs oracle (complex)
        assert abs(float(hx) - x) < 1e-6 and abs(float(hy) - y) < 1e-6
        hlo, hla = hwarp.utm_inverse(x, y, zone, south)
        assert abs(float(hlo) - lo) < 1e-11 and abs(float(hla) - la) < 1e-11
    assert owarp.epsg_to_utm(32611) == (11, False) and hwarp.epsg_to_utm(32734) == (34, True)
    with pytest.raises(ValueError):
        hwarp.epsg_to_utm(4326)


def _case():
    # an EMIT-like ortho grid: 0.000542 deg pixels near 34.1 N, 118.3 W (UTM zone 11), and an S2 10 m tile around it
    src_gt = (-118.60, 0.000542232520256367, 0.0, 34.45, 0.0, -0.000542232520256367)
    Hs, Ws = 1200, 1100
    s2 = hwarp.S2Grid(epsg=32611, x0=300000.0, y0=3900000.0, dx=10.0, dy=10.0, width=10980, height=10980)
    return src_gt, (Hs, Ws), s2


def test_compute_te_matches_oracle_and_snaps_inwards():
    src_gt, (Hs, Ws), s2 = _case()
    te = hwarp.compute_te(hwarp.bounds_of(src_gt, Ws, Hs), s2)
    ote = owarp.compute_te(owarp.bounds_of(src_gt, Ws, Hs), s2.bounds, (s2.x0, s2.y0), 11, False)
    assert te == pytest.approx(ote, abs=1e-6)
    l, b, r, t = te
    for v, o in ((l, s2.x0), (r, s2.x0), (t, s2.y0), (b, s2.y0)):
        assert abs((v - o) / 60.0 - round((v - o) / 60.0)) < 1e-9          # on the 60 m lattice anchored at the S2 origin
    sl, sb, sr, st = s2.bounds
    assert sl <= l < r <= sr and sb <= b < t <= st                         # inside the S2 tile
    dst_gt, (rows, cols), rec = hwarp.target_grid(src_gt, (Hs, Ws), s2)
    assert (rec["cols"], rec["rows"]) == (cols, rows) and cols == round((r - l) / 60) and rows == round((t - b) / 60)
    assert dst_gt == pytest.approx((l, 60.0, 0.0, t, 0.0, -60.0))
    with pytest.raises(ValueError):                                         # no overlap
        hwarp.compute_te((10.0, 10.0, 10.5, 10.5), s2)
    with pytest.raises(ValueError):                                         # 60 m is not a multiple of a 7 m grid
        hwarp.target_grid(src_gt, (Hs, Ws), hwarp.S2Grid(32611, 300000.0, 3900000.0, 7.0, 7.0, 100, 100))


def test_scales_and_coordinates_host_vs_oracle():
    src_gt, (Hs, Ws), s2 = _case()
    dst_gt, shape, _ = hwarp.target_grid(src_gt, (Hs, Ws), s2)
    hs = hwarp.warp_scales(dst_gt, src_gt, shape, 11, False)
    os_ = owarp.warp_scales(dst_gt, src_gt, shape[0], shape[1], 11, False, True)
    assert hs == pytest.approx(os_, rel=1e-12)
    # at 34 N a 0.000542 deg pixel is ~50 m wide and ~60 m tall (and the UTM box is slightly rotated): the 60 m grid is coarser in x only
    assert 0.80 < hs[0] < 0.86 and 0.95 < hs[1] < 1.02
    c, r = np.array([0.5, 100.5, shape[1] - 0.5]), np.array([0.5, 50.5, shape[0] - 0.5])
    hx, hy = hwarp.dst_to_src(c, r, dst_gt, src_gt, 11, False)
    for k in range(3):
        ox, oy = owarp.dst_to_src(c[k] - 0.5, r[k] - 0.5, dst_gt, src_gt, 11, False, True)
        assert abs(hx[k] - ox) < 1e-8 and abs(hy[k] - oy) < 1e-8


def test_kernel_weights():
    # cubic convolution (a = -0.5): interpolating, partition of unity, support 2
    assert owarp.cubic_weight(0.0) == 1.0 and owarp.cubic_weight(1.0) == 0.0 and owarp.cubic_weight(2.0) == 0.0
    assert owarp.cubic_weight(2.5) == 0.0 and owarp.cubic_weight(-0.5) == owarp.cubic_weight(0.5) == pytest.approx(0.5625)
    for d in np.linspace(0, 1, 11):
        assert sum(owarp.cubic_weight(i - d) for i in range(-1, 3)) == pytest.approx(1.0, abs=1e-14)
        assert sum(owarp.bilinear_weight(i - d) for i in range(0, 2)) == pytest.approx(1.0, abs=1e-14)


def test_oracle_warp_identity_and_nodata():
    rng = np.random.default_rng(1)
    src = rng.random((9, 11, 3)).astype(np.float32)
    gt = (100.0, 2.0, 0.0, 50.0, 0.0, -2.0)
    out = owarp.warp(src, gt, gt, 9, 11, utm=False, nodata=-9999.0)
    assert np.array_equal(out, src)                                          # same grid: taps (0, 1, 0, 0) -> a copy
    src[4, 5, :] = -9999.0                                                   # a hole: renormalised around it, hole itself
    out = owarp.warp(src, gt, gt, 9, 11, utm=False, nodata=-9999.0)          # keeps nodata (weight 1 tap is invalid)
    assert np.all(out[4, 5] == -9999.0) and np.array_equal(out[0, 0], src[0, 0])
    half = (100.0, 1.0, 0.0, 50.0, 0.0, -1.0)                                # 2x finer grid: still inside [min, max] + overshoot
    up = owarp.warp(src, gt, half, 18, 22, utm=False, nodata=-9999.0)
    assert up.shape == (18, 22, 3) and np.isfinite(up).all()
    far = (1000.0, 2.0, 0.0, 50.0, 0.0, -2.0)                                # no overlap: all nodata
    assert np.all(owarp.warp(src, gt, far, 4, 4, utm=False, nodata=-9999.0) == -9999.0)


def test_target_grid_southern_hemisphere_and_match_res():
    """Zone 34 south (Cape Town): the extent arithmetic goes through the false northing; a 20 m step on a 10 m S2 grid."""
    src_gt = (18.30, 0.000542232520256367, 0.0, -33.80, 0.0, -0.000542232520256367)
    s2 = hwarp.S2Grid(epsg=32734, x0=199980.0, y0=6300040.0, dx=10.0, dy=10.0, width=10980, height=10980)
    te = hwarp.compute_te(hwarp.bounds_of(src_gt, 900, 800), s2, 20.0, 20.0)
    ote = owarp.compute_te(owarp.bounds_of(src_gt, 900, 800), s2.bounds, (s2.x0, s2.y0), 34, True, 20.0, 20.0)
    assert te == pytest.approx(ote, abs=1e-6)
    dst_gt, (rows, cols), rec = hwarp.target_grid(src_gt, (800, 900), s2, 20.0, 20.0)
    assert rows > 0 and cols > 0 and dst_gt[1] == pytest.approx(20.0) and dst_gt[5] == pytest.approx(-20.0)
    assert 6.0e6 < rec["bottom"] < rec["top"] < 6.4e6                     # northings below 10 000 km: southern hemisphere
    # the centre of the target grid maps back into the source
    sx, sy = hwarp.dst_to_src(np.array([cols / 2.0]), np.array([rows / 2.0]), dst_gt, src_gt, 34, True)
    assert 0 < sx[0] < 900 and 0 < sy[0] < 800
    ox, oy = owarp.dst_to_src(cols / 2.0 - 0.5, rows / 2.0 - 0.5, dst_gt, src_gt, 34, True, True)
    assert abs(sx[0] - ox) < 1e-8 and abs(sy[0] - oy) < 1e-8

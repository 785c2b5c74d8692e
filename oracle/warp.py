"""Oracle for the general grid warp: the WGS-84 ortho cube -> Sentinel-2 UTM 60 m grid with
cubic resampling that ``nc_to_envi`` delegates to ``gdalwarp`` (EMIT_data/emit_proj.py:876-940), and the
snapped-extent arithmetic around it (``_bounds_to_out_crs`` :309-324, ``_intersect`` :326-332, ``_compute_te``
:354-382).  Test infrastructure — see oracle/__init__.py.

PARITY UNPINNED against GDAL / PROJ: neither is installed here (no gdalwarp, rasterio, pyproj), there is no
network, and the reference holds no vector at that boundary.  What is restated is the PUBLISHED algorithm:

* PROJ's transverse Mercator for EPSG:326xx / 327xx (``+proj=utm``): the Krueger series in the form and to the
  order (n^6) of Karney, "Transverse Mercator with an accuracy of a few nanometers" (J. Geodesy 85, 2011),
  eqs. 7-11, 35, 36 — the "exact" (Poder/Engsager) algorithm of PROJ agrees with it to well below a micrometre
  inside a UTM zone.  Written here with COMPLEX arithmetic (zeta = xi + i eta) so that it shares no code with the
  real-arithmetic implementations of the package (host numpy, device fp64).  Pinned by known answers: the
  false-easting limits of a zone on the equator (166 021.443 m / 833 978.557 m), the meridian arc to 45 deg N,
  and a landmark quoted in metres (tests/test_warp_oracle.py).
* GDAL's warp kernel for ``-r cubic`` with ``-srcnodata`` (gdalwarpkernel.cpp, general case with per-band
  validity masks): destination pixel centre -> source pixel coordinates through the exact transformer
  (``-et 0``; gdalwarp's default approximates it to 0.125 px), cubic convolution weights with a = -0.5
  (``GWKCubic``), filter radius 2 widened to ceil(2 / scale) with the argument multiplied by ``scale`` where the
  destination is coarser than the source (scale = destination size / source window size, one value per axis for
  the whole warp), taps outside the source or equal to nodata skipped, the sum divided by the accumulated
  weight, destination left at nodata when the centre falls outside the source or the accumulated weight is
  below 1e-6.  NaN is an ordinary value (it is not the nodata) and propagates, except under a weight of exactly
  zero (a tap outside the filter's support contributes nothing).
"""
from __future__ import annotations

import cmath
import math

import numpy as np

WGS84_A = 6378137.0
WGS84_F = 1.0 / 298.257223563
UTM_K0 = 0.9996


def _series():
    n = WGS84_F / (2.0 - WGS84_F)
    n2, n3, n4, n5, n6 = n ** 2, n ** 3, n ** 4, n ** 5, n ** 6
    A = WGS84_A / (1.0 + n) * (1.0 + n2 / 4.0 + n4 / 64.0 + n6 / 256.0)
    alpha = [
        n / 2 - 2 * n2 / 3 + 5 * n3 / 16 + 41 * n4 / 180 - 127 * n5 / 288 + 7891 * n6 / 37800,
        13 * n2 / 48 - 3 * n3 / 5 + 557 * n4 / 1440 + 281 * n5 / 630 - 1983433 * n6 / 1935360,
        61 * n3 / 240 - 103 * n4 / 140 + 15061 * n5 / 26880 + 167603 * n6 / 181440,
        49561 * n4 / 161280 - 179 * n5 / 168 + 6601661 * n6 / 7257600,
        34729 * n5 / 80640 - 3418889 * n6 / 1995840,
        212378941 * n6 / 319334400,
    ]
    beta = [
        n / 2 - 2 * n2 / 3 + 37 * n3 / 96 - n4 / 360 - 81 * n5 / 512 + 96199 * n6 / 604800,
        n2 / 48 + n3 / 15 - 437 * n4 / 1440 + 46 * n5 / 105 - 1118711 * n6 / 3870720,
        17 * n3 / 480 - 37 * n4 / 840 - 209 * n5 / 4480 + 5569 * n6 / 90720,
        4397 * n4 / 161280 - 11 * n5 / 504 - 830251 * n6 / 7257600,
        4583 * n5 / 161280 - 108847 * n6 / 3991680,
        20648693 * n6 / 638668800,
    ]
    return A, alpha, beta


_A, _ALPHA, _BETA = _series()
_E = math.sqrt(WGS84_F * (2.0 - WGS84_F))


def utm_lon0(zone: int) -> float:
    return -183.0 + 6.0 * zone


def _taup(tau):
    sigma = math.sinh(_E * math.atanh(_E * tau / math.hypot(1.0, tau)))
    return tau * math.hypot(1.0, sigma) - sigma * math.hypot(1.0, tau)


def utm_forward(lon, lat, zone: int, south: bool = False):
    """(lon, lat) degrees -> (easting, northing) metres.  Scalar."""
    lam = math.radians(lon - utm_lon0(zone))
    tau = math.tan(math.radians(lat))
    tp = _taup(tau)
    zeta_p = complex(math.atan2(tp, math.cos(lam)), math.asinh(math.sin(lam) / math.hypot(tp, math.cos(lam))))
    zeta = zeta_p + sum(a * cmath.sin(2 * (j + 1) * zeta_p) for j, a in enumerate(_ALPHA))
    x = UTM_K0 * _A * zeta.imag + 500000.0
    y = UTM_K0 * _A * zeta.real + (10000000.0 if south else 0.0)
    return x, y


def utm_inverse(x, y, zone: int, south: bool = False):
    """(easting, northing) metres -> (lon, lat) degrees.  Scalar."""
    zeta = complex((y - (10000000.0 if south else 0.0)) / (UTM_K0 * _A), (x - 500000.0) / (UTM_K0 * _A))
    zeta_p = zeta - sum(b * cmath.sin(2 * (j + 1) * zeta) for j, b in enumerate(_BETA))
    xi, eta = zeta_p.real, zeta_p.imag
    tp = math.sin(xi) / math.hypot(math.sinh(eta), math.cos(xi))
    lam = math.atan2(math.sinh(eta), math.cos(xi))
    tau = tp
    for _ in range(5):                                   # Newton on tau'(tau) = tp (Karney eqs. 19-21)
        ti = _taup(tau)
        e2 = _E * _E
        dtau = (tp - ti) / math.hypot(1.0, ti) * (1.0 + (1.0 - e2) * tau * tau) / ((1.0 - e2) * math.hypot(1.0, tau))
        tau += dtau
        if abs(dtau) < 1e-15 * max(1.0, abs(tau)):
            break
    return math.degrees(lam) + utm_lon0(zone), math.degrees(math.atan(tau))


def epsg_to_utm(epsg: int):
    """EPSG:326zz -> (zone, north), EPSG:327zz -> (zone, south)."""
    if 32601 <= epsg <= 32660:
        return epsg - 32600, False
    if 32701 <= epsg <= 32760:
        return epsg - 32700, True
    raise ValueError(f"EPSG:{epsg} is not a WGS-84 UTM zone")


# ------------------------------------------------------------------ extent snapping (emit_proj.py:309-382)
def bounds_of(gt, width, height):
    """rasterio's ds.bounds for a north-up geotransform (x_ul, x_res, 0, y_ul, 0, -y_res)."""
    left, top = gt[0], gt[3]
    right, bottom = gt[0] + width * gt[1], gt[3] + height * gt[5]
    return left, bottom, right, top


def bounds_to_utm(src_bounds, zone, south):            # _bounds_to_out_crs :309-324 (the 4 corners only)
    l, b, r, t = src_bounds
    pts = [utm_forward(x, y, zone, south) for x, y in ((l, b), (l, t), (r, b), (r, t))]
    xs, ys = [p[0] for p in pts], [p[1] for p in pts]
    return min(xs), min(ys), max(xs), max(ys)


def intersect(a, b):                                    # _intersect :326-332
    l, bb, r, t = max(a[0], b[0]), max(a[1], b[1]), min(a[2], b[2]), min(a[3], b[3])
    if r <= l or t <= bb:
        return None
    return l, bb, r, t


def compute_te(src_bounds, s2_te_exact, s2_origin_xy, zone, south, xres=60.0, yres=60.0):   # _compute_te :354-382
    inter = intersect(bounds_to_utm(src_bounds, zone, south), tuple(map(float, s2_te_exact)))
    if inter is None:
        raise ValueError("No overlap between EMIT source bounds and S2 extent in out_crs.")
    il, ib, ir, it = inter
    x0, y0 = map(float, s2_origin_xy)
    eps = 1e-9
    left = x0 + math.ceil(((il - x0) / xres) - eps) * xres
    right = x0 + math.floor(((ir - x0) / xres) + eps) * xres
    top = y0 - math.ceil(((y0 - it) / yres) - eps) * yres
    bottom = y0 - math.floor(((y0 - ib) / yres) + eps) * yres
    if right <= left or top <= bottom:
        raise ValueError(f"Snapped TE is invalid: {(left, bottom, right, top)}")
    return left, bottom, right, top


# ------------------------------------------------------------------ the warp kernel
def cubic_weight(x):                                    # GWKCubic, a = -0.5
    ax = abs(x)
    if ax <= 1.0:
        return (1.5 * ax - 2.5) * ax * ax + 1.0
    if ax <= 2.0:
        return ((-0.5 * ax + 2.5) * ax - 4.0) * ax + 2.0
    return 0.0


def bilinear_weight(x):                                 # GWKBilinear
    ax = abs(x)
    return 1.0 - ax if ax <= 1.0 else 0.0


def dst_to_src(col, row, dst_gt, src_gt, zone, south, utm):
    """Centre of destination pixel (col, row) -> source pixel coordinates (corner convention: the centre of source
    pixel (i, j) is (i + 0.5, j + 0.5)).  utm=False: both grids share one CRS (affine only)."""
    X = dst_gt[0] + (col + 0.5) * dst_gt[1] + (row + 0.5) * dst_gt[2]
    Y = dst_gt[3] + (col + 0.5) * dst_gt[4] + (row + 0.5) * dst_gt[5]
    if utm:
        X, Y = utm_inverse(X, Y, zone, south)
    det = src_gt[1] * src_gt[5] - src_gt[2] * src_gt[4]
    dx, dy = X - src_gt[0], Y - src_gt[3]
    return (src_gt[5] * dx - src_gt[2] * dy) / det, (-src_gt[4] * dx + src_gt[1] * dy) / det


def warp_scales(dst_gt, src_gt, Hd, Wd, zone, south, utm, npts=21):
    """(xscale, yscale) = destination size / extent of the source window, the window being the bounding box of the
    destination's edges (npts + 1 samples per edge, GDAL's ComputeSourceWindow) in source pixel coordinates."""
    sx, sy = [], []
    for k in range(npts + 1):
        t = k / npts
        for c, r in ((t * Wd, 0.0), (t * Wd, float(Hd)), (0.0, t * Hd), (float(Wd), t * Hd)):
            x, y = dst_to_src(c - 0.5, r - 0.5, dst_gt, src_gt, zone, south, utm)     # edge points, not centres
            sx.append(x)
            sy.append(y)
    return Wd / (max(sx) - min(sx)), Hd / (max(sy) - min(sy))


def warp(src, src_gt, dst_gt, Hd, Wd, *, zone=0, south=False, utm=True, nodata=None, dst_nodata=None,
         kernel="cubic", scales=None):
    """src (Hs, Ws, B) float32 -> (Hd, Wd, B) float32.  Pure-Python loops over destination pixels: small cases only."""
    src = np.asarray(src, dtype=np.float32)
    Hs, Ws, B = src.shape
    fill = np.float32(dst_nodata if dst_nodata is not None else (nodata if nodata is not None else 0.0))
    out = np.full((Hd, Wd, B), fill, dtype=np.float32)
    if kernel in ("nearest", "average"):
        return _warp_point(src, out, src_gt, dst_gt, Hd, Wd, zone, south, utm, nodata, kernel)
    xs, ys = scales if scales is not None else warp_scales(dst_gt, src_gt, Hd, Wd, zone, south, utm)
    wfun, r0 = (cubic_weight, 2) if kernel == "cubic" else (bilinear_weight, 1)
    fx, fy = min(xs, 1.0), min(ys, 1.0)
    rx, ry = (r0 if xs >= 1.0 else int(math.ceil(r0 / xs))), (r0 if ys >= 1.0 else int(math.ceil(r0 / ys)))
    s64 = src.astype(np.float64)
    for row in range(Hd):
        for col in range(Wd):
            sx, sy = dst_to_src(col, row, dst_gt, src_gt, zone, south, utm)
            if not (0.0 <= sx < Ws and 0.0 <= sy < Hs):            # centre outside the source: stays nodata
                continue
            ix, iy = int(math.floor(sx - 0.5)), int(math.floor(sy - 0.5))
            ddx, ddy = sx - 0.5 - ix, sy - 0.5 - iy
            acc = np.zeros(B)
            wsum = np.zeros(B)
            for j in range(1 - ry, ry + 1):
                yy = iy + j
                if yy < 0 or yy >= Hs:
                    continue
                wy = wfun((j - ddy) * fy)
                for i in range(1 - rx, rx + 1):
                    xx = ix + i
                    if xx < 0 or xx >= Ws:
                        continue
                    w = np.float64(np.float32(wy * wfun((i - ddx) * fx)))   # the kernels carry the tap weight in fp32
                    if w == 0.0:                                            # outside the filter's support: no contribution,
                        continue                                            # not even of a NaN
                    v = s64[yy, xx]
                    ok = np.ones(B, bool) if nodata is None else (src[yy, xx] != np.float32(nodata))
                    acc += np.where(ok, w * v, 0.0)
                    wsum += np.where(ok, w, 0.0)
            good = wsum >= 1e-6
            with np.errstate(invalid="ignore", divide="ignore"):
                out[row, col, good] = (acc[good] / wsum[good]).astype(np.float32)
    return out


def _cover(i, i0, i1, lo, hi):
    """Fraction of source pixel ``i`` of [i0, i1) the footprint [lo, hi] covers along one axis (GDAL's COMPUTE_WEIGHT)."""
    if i == i0:
        return 1.0 if i0 + 1 == i1 else 1.0 - (lo - i0)
    if i + 1 == i1:
        return 1.0 - (i1 - hi)
    return 1.0


def _warp_point(src, out, src_gt, dst_gt, Hd, Wd, zone, south, utm, nodata, kernel):
    """GWKNearest / GWKAverageOrMode of gdalwarpkernel.cpp, restated (parity with GDAL unpinned): nearest = the source
    pixel floor(x + 1e-10), floor(y + 1e-10) under the destination centre; average = all source pixels of the box between
    the transformed top-left and bottom-right corners of the destination pixel (clipped to the source), weighted by the
    covered fraction per axis, nodata skipped per band, float64 sums."""
    Hs, Ws, B = src.shape
    nd = None if nodata is None else np.float32(nodata)
    for row in range(Hd):
        for col in range(Wd):
            if kernel == "nearest":
                sx, sy = dst_to_src(col, row, dst_gt, src_gt, zone, south, utm)
                if not (0.0 <= sx <= Ws and 0.0 <= sy <= Hs):
                    continue
                ix, iy = int(math.floor(sx + 1e-10)), int(math.floor(sy + 1e-10))
                ix, iy = min(ix, Ws - 1), min(iy, Hs - 1)
                v = src[iy, ix]
                ok = np.ones(B, bool) if nd is None else (v != nd)
                out[row, col, ok] = v[ok]
                continue
            ax, ay = dst_to_src(col - 0.5, row - 0.5, dst_gt, src_gt, zone, south, utm)      # corners, not centres
            bx, by = dst_to_src(col + 0.5, row + 0.5, dst_gt, src_gt, zone, south, utm)
            xmin, xmax, ymin, ymax = min(ax, bx), max(ax, bx), min(ay, by), max(ay, by)
            if not (xmax > 0.0 and ymax > 0.0 and xmin < Ws and ymin < Hs):
                continue
            xmin, ymin, xmax, ymax = max(xmin, 0.0), max(ymin, 0.0), min(xmax, float(Ws)), min(ymax, float(Hs))
            x0, x1 = int(math.floor(xmin + 1e-10)), int(math.ceil(xmax - 1e-10))
            y0, y1 = int(math.floor(ymin + 1e-10)), int(math.ceil(ymax - 1e-10))
            if x0 == x1 and x1 < Ws:
                x1 += 1
            if y0 == y1 and y1 < Hs:
                y1 += 1
            tot = np.zeros(B)
            wsum = np.zeros(B)
            for y in range(y0, y1):
                wy = _cover(y, y0, y1, ymin, ymax)
                for x in range(x0, x1):
                    w = wy * _cover(x, x0, x1, xmin, xmax)
                    v = src[y, x]
                    ok = np.ones(B, bool) if nd is None else (v != nd)
                    tot += np.where(ok, v.astype(np.float64) * w, 0.0)
                    wsum += np.where(ok, w, 0.0)
            good = wsum > 0.0
            with np.errstate(invalid="ignore", divide="ignore"):
                out[row, col, good] = (tot[good] / wsum[good]).astype(np.float32)
    return out

"""Pin the GDAL boundary the moment GDAL is available:

    python tests/golden/make_golden_gdal.py     ->  tests/golden/gdal_warp.npz

Needs the `gdalwarp` CLI (what the reference shells out to, EMIT_data/emit_proj.py:876-940) and the `osgeo.gdal`
Python bindings to write / read the rasters; `rasterio` adds the notebook's `rasterio.warp.reproject` cases
(Pairs_EMIT_S2_demo-2.ipynb cell 73 `downsample_s2_to_grid` "average" / `reproject_stack_to_grid` "bilinear";
s2_data/s2_utils.py:546-558).  None of them is installed in the build image (no wheel, no network), so the file is
absent, tests/test_gdal_pin.py SKIPS with "parity unpinned", and oracle/warp.py / oracle/resample.py (restatements of
the published GDAL / PROJ algorithms) stay the checkers, pinned only where known answers exist (UTM known answers, the
cubic kernel's documented values, identity warps).

Two gdalwarp goldens per case: the reference's EXACT command line (default error threshold 0.125 px: an approximate
transformer) and the same with `-et 0` (exact transformer — what hsr_warp_f32 implements)."""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gdal_warp.npz")
NODATA = -9999.0


def main():
    try:
        from osgeo import gdal, osr
    except ImportError:
        print("osgeo.gdal is not installed: nothing generated; parity with gdalwarp stays unpinned")
        return 1
    if shutil.which("gdalwarp") is None:
        print("gdalwarp CLI not on PATH: nothing generated; parity with gdalwarp stays unpinned")
        return 1
    from hsr_b200.EMIT_data import warp as hwarp

    rng = np.random.default_rng(20240821)
    save = {"gdal_version": np.array(gdal.VersionInfo("RELEASE_NAME"))}
    tmp = tempfile.mkdtemp()
    # an EMIT-like WGS-84 ortho grid near 34 N, 118 W (UTM 11 N) and a small S2 10 m grid around it
    src_gt = (-118.30, 0.000542232520256367, 0.0, 34.20, 0.0, -0.000542232520256367)
    Hs, Ws, B = 96, 80, 5
    yy, xx = np.mgrid[0:Hs, 0:Ws]
    cube = np.stack([0.3 + 0.2 * np.sin(0.11 * xx + 0.3 * b) * np.cos(0.07 * yy) + 0.02 * rng.random((Hs, Ws))
                     for b in range(B)], axis=-1).astype(np.float32)
    cube[:6, :, :] = NODATA                                                    # outside-swath margin
    cube[40:43, 30:34, :] = NODATA                                             # a hole
    cube[60, 50, 2] = NODATA                                                   # nodata in ONE band only
    s2 = hwarp.S2Grid(epsg=32611, x0=380000.0, y0=3790000.0, dx=10.0, dy=10.0, width=1200, height=1200)
    dst_gt, (Hd, Wd), rec = hwarp.target_grid(src_gt, (Hs, Ws), s2)
    drv = gdal.GetDriverByName("ENVI")
    src_path = os.path.join(tmp, "src")
    ds = drv.Create(src_path, Ws, Hs, B, gdal.GDT_Float32, options=["INTERLEAVE=BIL"])
    ds.SetGeoTransform(src_gt)
    srs = osr.SpatialReference()
    srs.ImportFromEPSG(4326)
    ds.SetProjection(srs.ExportToWkt())
    for b in range(B):
        ds.GetRasterBand(b + 1).WriteArray(cube[..., b])
    ds = None
    for tag, extra in (("ref", []), ("et0", ["-et", "0"])):
        dst_path = os.path.join(tmp, f"dst_{tag}")
        cmd = ["gdalwarp", "-overwrite", "--config", "GDAL_CACHEMAX", "2048", "-t_srs", "EPSG:32611",
               "-te", str(rec["left"]), str(rec["bottom"]), str(rec["right"]), str(rec["top"]),
               "-ts", str(Wd), str(Hd), "-srcnodata", str(NODATA), "-dstnodata", str(NODATA),
               "-multi", "-wo", "NUM_THREADS=ALL_CPUS", "-wm", "4096"] + extra + ["-r", "cubic", "-of", "ENVI", src_path, dst_path]
        subprocess.run(cmd, check=True, capture_output=True)
        out = gdal.Open(dst_path).ReadAsArray()                                 # (B, Hd, Wd)
        save[f"warp_{tag}"] = np.transpose(out, (1, 2, 0)).astype(np.float32)
    save.update({"src": cube, "src_gt": np.array(src_gt), "dst_gt": np.array(dst_gt), "dst_shape": np.array([Hd, Wd]),
                 "s2": np.array([s2.epsg, s2.x0, s2.y0, s2.dx, s2.dy, s2.width, s2.height], dtype=np.float64)})
    try:
        import rasterio  # noqa: F401
        from rasterio.crs import CRS
        from rasterio.transform import Affine
        from rasterio.warp import Resampling, reproject

        # notebook cell 73: S2 10 m stack -> EMIT 60 m grid, "average"; 60 m -> 10 m, "bilinear" (snapped, factor 6)
        fine = (rng.random((3, 120, 132)) * 255).astype(np.uint8)
        t10 = Affine(10.0, 0.0, 380000.0, 0.0, -10.0, 3790000.0)
        t60 = Affine(60.0, 0.0, 380000.0, 0.0, -60.0, 3790000.0)
        crs = CRS.from_epsg(32611)
        coarse = np.zeros((3, 20, 22), np.float32)
        for k in range(3):
            reproject(fine[k].astype(np.float32), coarse[k], src_transform=t10, src_crs=crs, dst_transform=t60, dst_crs=crs,
                      resampling=Resampling.average)
        up = np.zeros((3, 120, 132), np.float32)
        for k in range(3):
            reproject(coarse[k], up[k], src_transform=t60, src_crs=crs, dst_transform=t10, dst_crs=crs,
                      resampling=Resampling.bilinear)
        save.update({"rio_fine": fine, "rio_average": coarse, "rio_bilinear": up})
        # s2_data/s2_utils.py:546-574: 20 m bands onto the 10 m blue grid, "nearest" and "bilinear"; and "average" /
        # "nearest" onto a grid that is NOT snapped (shifted origin, non-integer ratio) -> the general warp kernel
        b20 = (rng.random((60, 66)) * 4000).astype(np.float32)
        t20 = Affine(20.0, 0.0, 380000.0, 0.0, -20.0, 3790000.0)
        tsh = Affine(50.0, 0.0, 380007.0, 0.0, -50.0, 3789990.0)
        for name, st, ssrc, dt, shape in (("20to10", t20, b20, t10, (120, 132)), ("10toshift", t10, fine[0].astype(np.float32), tsh, (22, 24))):
            for rs in ("nearest", "bilinear", "average"):
                dst = np.zeros(shape, np.float32)
                reproject(ssrc, dst, src_transform=st, src_crs=crs, dst_transform=dt, dst_crs=crs,
                          resampling=getattr(Resampling, rs))
                save[f"rio_{name}_{rs}"] = dst
            save[f"rio_{name}_src"] = ssrc
            save[f"rio_{name}_gts"] = np.array([[st.c, st.a, st.b, st.f, st.d, st.e], [dt.c, dt.a, dt.b, dt.f, dt.d, dt.e]])
    except ImportError:
        print("rasterio not installed: reproject cases skipped")
    np.savez_compressed(OUT, **save)
    shutil.rmtree(tmp, ignore_errors=True)
    print("wrote", OUT, "GDAL", save["gdal_version"])
    return 0


if __name__ == "__main__":
    sys.exit(main())

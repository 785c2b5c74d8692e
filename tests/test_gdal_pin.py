"""The pin for the GDAL boundary: oracle/warp.py, oracle/resample.py and the CUDA warp against
outputs of the REAL gdalwarp / rasterio (tests/golden/gdal_warp.npz, produced by tests/golden/make_golden_gdal.py
wherever GDAL is installed).  Absent in the build image -> these tests SKIP with "parity unpinned"."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gdal_warp.npz")
needs_gdal_golden = pytest.mark.skipif(
    not os.path.exists(GOLD),
    reason="parity unpinned: tests/golden/gdal_warp.npz absent — run tests/golden/make_golden_gdal.py where GDAL is installed")
NODATA = -9999.0


def _compare(got, want, rtol, atol, frac_edge=0.0):
    """Equal nodata footprint (up to `frac_edge` of the pixels, on the footprint's edge) and close values elsewhere."""
    gn, wn = got == NODATA, want == NODATA
    assert (gn != wn).mean() <= frac_edge
    ok = ~gn & ~wn
    assert ok.any() and np.allclose(got[ok], want[ok], rtol=rtol, atol=atol)


@needs_gdal_golden
def test_oracle_warp_equals_gdalwarp():
    from hsr_b200.EMIT_data import warp as hwarp
    from oracle import warp as owarp

    g = np.load(GOLD)
    Hd, Wd = [int(v) for v in g["dst_shape"]]
    scales = hwarp.warp_scales(tuple(g["dst_gt"]), tuple(g["src_gt"]), (Hd, Wd), 11, False)
    out = owarp.warp(g["src"], tuple(g["src_gt"]), tuple(g["dst_gt"]), Hd, Wd, zone=11, utm=True, nodata=NODATA, scales=scales)
    _compare(out, g["warp_et0"], rtol=1e-5, atol=1e-6)                      # exact transformer: the tight bar
    _compare(out, g["warp_ref"], rtol=0, atol=2e-2, frac_edge=0.02)         # the reference's flags: 0.125 px transformer error
    if "rio_average" in g.files:
        from oracle import resample as ores

        np.testing.assert_allclose(ores.downsample_to_grid(g["rio_fine"], 6), g["rio_average"], rtol=1e-6, atol=1e-4)
        np.testing.assert_allclose(ores.upsample_to_grid(g["rio_average"], 6), g["rio_bilinear"], rtol=1e-5, atol=1e-4)
    for name in ("20to10", "10toshift"):
        if f"rio_{name}_src" not in g.files:
            continue
        sgt, dgt = [tuple(v) for v in g[f"rio_{name}_gts"]]
        for rs in ("nearest", "bilinear", "average"):
            want = g[f"rio_{name}_{rs}"]
            scales = hwarp.warp_scales(dgt, sgt, want.shape)
            got = owarp.warp(g[f"rio_{name}_src"][..., None], sgt, dgt, want.shape[0], want.shape[1], utm=False, nodata=None,
                             kernel=rs, scales=scales)[..., 0]
            np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-3, err_msg=f"{name} {rs}")


@needs_gdal_golden
@pytest.mark.gpu
def test_cuda_warp_equals_gdalwarp():
    import torch

    from hsr_b200 import kernels
    from hsr_b200.EMIT_data import warp as hwarp

    g = np.load(GOLD)
    Hd, Wd = [int(v) for v in g["dst_shape"]]
    scales = hwarp.warp_scales(tuple(g["dst_gt"]), tuple(g["src_gt"]), (Hd, Wd), 11, False)
    out = kernels.warp(torch.from_numpy(g["src"]).cuda(), tuple(g["src_gt"]), tuple(g["dst_gt"]), (Hd, Wd), utm_zone=11,
                       scales=scales, kernel="cubic", nodata=NODATA).cpu().numpy()
    _compare(out, g["warp_et0"], rtol=1e-5, atol=1e-6)
    if "rio_average" in g.files:
        avg = kernels.block_average(torch.from_numpy(g["rio_fine"]).cuda(), 6).cpu().numpy()
        np.testing.assert_allclose(avg, g["rio_average"], rtol=1e-6, atol=1e-4)
        up = kernels.bilinear_upsample(torch.from_numpy(g["rio_average"]).cuda(), 6).cpu().numpy()
        np.testing.assert_allclose(up, g["rio_bilinear"], rtol=1e-5, atol=1e-4)
    for name in ("20to10", "10toshift"):
        if f"rio_{name}_src" not in g.files:
            continue
        sgt, dgt = [tuple(v) for v in g[f"rio_{name}_gts"]]
        for rs in ("nearest", "bilinear", "average"):
            want = g[f"rio_{name}_{rs}"]
            got = hwarp.warp_to_grid(g[f"rio_{name}_src"], sgt, dgt, want.shape, kernel=rs, nodata=None)
            np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-3, err_msg=f"{name} {rs}")

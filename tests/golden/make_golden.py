"""Generate tests/golden/*.npz by running the REFERENCE'S OWN functions (from /root/reference)
on small seeded inputs.  Run in the build container (the reference is not available on the GPU
box):      python tests/golden/make_golden.py

The reference functions are loaded by oracle/ref_loader.py (AST extraction of the pure-numpy
FunctionDefs; nothing is copied).  Inputs and outputs are stored together so the tests never
depend on RNG stability.
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from hsr_b200 import synthetic  # noqa: E402
from hsr_b200.s2_emit.srf import synthetic_s2_srf  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
warnings.simplefilter("ignore", DeprecationWarning)


def main():
    ref = ref_loader.load()
    rng = np.random.default_rng(20240819)
    w = synthetic.emit_wavelengths()
    good = synthetic.good_band_mask(w)

    # ---- GLT gather: reference apply_glt (EMIT_data/emit_tools.py:153-181), in-range GLT with holes
    Hr, Wr, B = 11, 9, 285
    raw = synthetic.raw_cube_bits_np((Hr, Wr, B), seed=11, good=good)
    raw[2, 3, 17] = np.nan
    raw[5, 1, 200] = np.inf
    gx, gy = synthetic.rotation_glt(Hr, Wr, 25.0)
    holes = rng.random(gx.shape) < 0.05
    gx[holes] = 0
    glt = np.stack([gx, gy], axis=-1).astype(np.int64)
    ortho = ref.apply_glt(raw, glt)
    plane = ref.apply_glt(raw[..., 40], glt)            # 2-D input flavour (elev / LOC planes)
    np.savez_compressed(os.path.join(OUT, "glt_apply_glt.npz"), raw=raw, glt=glt, ortho=ortho, plane=plane)

    # ---- SRF synthesis: reference pseudo_s2_srf_integral / pseudo_s2_rgb (s2_emit/synth.py:9-58)
    cube = synthetic.raw_cube_spectra_np((7, 6, B), seed=5, good=good)
    cube[0, 0, :] = -9999.0          # a fill pixel integrates to ~ -9999 (no masking before SRF)
    cube[1, 2, 3] = np.nan           # NaN in a band with ZERO weight still poisons every band
    cube[3, 4, 30] = np.inf          # Inf under a zero weight -> NaN everywhere
    cube[4, 1, 38] = np.inf          # Inf inside B4's response only: B4 -> +Inf, every other band -> NaN
    cube[5, 2, 160] = -np.inf        # -Inf inside B11's response: B11 -> -Inf
    srf = synthetic_s2_srf()
    ps_good = ref.pseudo_s2_srf_integral(cube, w, srf, good)
    ps_all = ref.pseudo_s2_srf_integral(cube, w, srf, None)
    rgb = ref.pseudo_s2_rgb(ps_good)
    names = list(srf.keys())
    save = {"cube": cube, "emit_w": w, "good": good, "names": np.array(names), "rgb": rgb}
    for b in names:
        save[f"lam_{b}"], save[f"rsp_{b}"] = srf[b]
        save[f"none_good_{b}"] = np.array(ps_good[b] is None)
        save[f"none_all_{b}"] = np.array(ps_all[b] is None)
        if ps_good[b] is not None:
            save[f"out_good_{b}"] = ps_good[b]
        if ps_all[b] is not None:
            save[f"out_all_{b}"] = ps_all[b]
    np.savez_compressed(os.path.join(OUT, "srf_pseudo_s2.npz"), **save)

    # ---- polynomial apply: reference apply_poly_rgb (s2_emit/poly_regression.py:65-84)
    H, Wd = 24, 20
    img = rng.uniform(-0.2, 1.3, size=(H, Wd, 3)).astype(np.float32)
    img[0, 0, 1] = np.nan
    img[1, 1, 2] = np.inf
    mask = rng.random((H, Wd)) < 0.7
    coeffs2 = np.array([[-0.31, 1.12, 0.021], [-0.28, 1.07, 0.018], [0.4, 0.55, -0.03]])
    coeffs4 = rng.normal(0, 0.5, size=(3, 5))
    # np.polyfit per channel — the call at poly_regression.py:58-60 — on pixel-paired data
    yimg = (coeffs2[:, 0] * img.astype(np.float64) ** 2 + coeffs2[:, 1] * img + coeffs2[:, 2]
            + rng.normal(0, 0.01, size=img.shape)).astype(np.float32)
    keep = mask & np.isfinite(img).all(-1) & np.isfinite(yimg).all(-1)
    fit2 = np.stack([np.polyfit(img[..., c][keep].astype(np.float64), yimg[..., c][keep].astype(np.float64), 2)
                     for c in range(3)])
    fit4 = np.stack([np.polyfit(img[..., c][keep].astype(np.float64), yimg[..., c][keep].astype(np.float64), 4)
                     for c in range(3)])
    # identity branch of the reference's fit_ot_poly_rgb (< 200 samples, poly_regression.py:38-41)
    small_mask = np.zeros((H, Wd), bool)
    small_mask[:5, :5] = True
    ident = ref.fit_ot_poly_rgb(img, yimg, small_mask, deg=2)
    np.savez_compressed(
        os.path.join(OUT, "poly_apply_fit.npz"), img=img, yimg=yimg, mask=mask, coeffs2=coeffs2, coeffs4=coeffs4,
        out2_mask=ref.apply_poly_rgb(img, coeffs2, mask), out2_nomask=ref.apply_poly_rgb(img, coeffs2, None),
        out4_mask=ref.apply_poly_rgb(img, coeffs4, mask), fit2=fit2, fit4=fit4, ident=ident, small_mask=small_mask)
    # ---- shared percentile stretch: reference apply_shared_percentile_stretch (s2_emit/color.py:25-34)
    H, Wd = 37, 29                                     # 1073 px: odd, planes not 16-byte multiples
    simg = (rng.random((H, Wd, 3)) ** 2 * np.array([0.6, 0.35, 0.2])).astype(np.float32)
    simg[..., 1] = np.round(simg[..., 1] * 50) / 50    # heavy ties in one channel
    simg[3, 4, 0] = -0.0
    simg[5, 6, 2] = np.inf                             # finite percentiles, Inf clips to 1
    simg[7, 7, 0] = -np.inf
    smask = rng.random((H, Wd)) < 0.55
    smask[5, 6] = smask[7, 7] = smask[3, 4] = True
    with np.errstate(invalid="ignore"):
        sout = ref.apply_shared_percentile_stretch(simg, smask)
        sout_1_99 = ref.apply_shared_percentile_stretch(simg, smask, 1, 99.5)
    slim = np.stack([np.percentile(simg[..., c][smask], [2, 98]) for c in range(3)])
    tiny_mask = np.zeros((H, Wd), bool)
    tiny_mask[10, 10:13] = True                        # 3 samples: both percentiles between neighbours
    tiny = ref.apply_shared_percentile_stretch(simg, tiny_mask)
    nanimg = simg.copy()
    nanimg[20, 20, 1] = np.nan                         # a NaN inside the mask: that channel is all NaN
    nmask = smask.copy()
    nmask[20, 20] = True
    with np.errstate(invalid="ignore"):
        nout = ref.apply_shared_percentile_stretch(nanimg, nmask)
    np.savez_compressed(os.path.join(OUT, "color_stretch.npz"), img=simg, mask=smask, out=sout, out_1_995=sout_1_99,
                        limits=slim, tiny_mask=tiny_mask, tiny=tiny, nanimg=nanimg, nmask=nmask, nout=nout)
    # ---- OT-target fit: the reference's own fit_ot_poly_rgb (s2_emit/poly_regression.py:16-62) with its
    #      `ot.dist` / `ot.sinkhorn` calls resolved to oracle/ot.py (POT is absent: parity unpinned THERE only)
    from oracle import ot as oot
    H, Wd = 64, 57
    osrc = (rng.random((H, Wd, 3)) ** 1.5 * np.array([0.9, 0.7, 0.5])).astype(np.float32)
    oref = np.clip(0.8 * osrc.astype(np.float64) ** 2 + 0.15 * osrc + 0.02 + rng.normal(0, 0.02, osrc.shape), 0, 1)
    oref = oref.astype(np.float32)
    osrc[2, 3, 1] = np.nan                 # dropped from X_all only
    oref[5, 6, 0] = np.inf                 # dropped from Y_all only: X and Y are filtered independently (:35-36)
    omask = rng.random((H, Wd)) < 0.8
    omask[2, 3] = omask[5, 6] = True
    ofit = {}
    for deg, ns_, seed in ((2, 600, 0), (4, 600, 3), (2, 100000, 1)):     # the last: fewer rows than n_samples
        ofit[f"coeffs_d{deg}_n{ns_}_s{seed}"] = ref.fit_ot_poly_rgb(osrc, oref, omask, deg=deg, n_samples=ns_, seed=seed)
    small = np.zeros((H, Wd), bool)
    small[:10, :10] = True
    ofit["ident"] = ref.fit_ot_poly_rgb(osrc, oref, small, deg=3)
    Xs = osrc[omask].reshape(-1, 3).astype(np.float64)
    Xs = Xs[np.isfinite(Xs).all(1)][:500]
    Ys = oref[omask].reshape(-1, 3).astype(np.float64)
    Ys = Ys[np.isfinite(Ys).all(1)][100:530]
    ofit["X"], ofit["Y"] = Xs, Ys
    ofit["ybar"] = oot.barycentric_targets(Xs, Ys, 0.05, 300, 1e-6)
    ofit["ybar_12it"] = oot.barycentric_targets(Xs, Ys, 0.05, 12, 0.0)     # iteration cap, no convergence
    np.savez_compressed(os.path.join(OUT, "ot_fit.npz"), src=osrc, ref=oref, mask=omask, small=small, **ofit)
    # ---- tile validity + band subsample: reference is_black_mask / _subsample_bands_evenly (tiles_helpers/utils.py)
    tb, th, tw = 9, 21, 19
    tile = rng.uniform(0.0, 0.6, size=(tb, th, tw)).astype(np.float32)
    tile[:, 0, 0] = -9999.0                                  # nodata in every band
    tile[:, 0, 1] = -9999.0
    tile[4, 0, 1] = 0.3                                      # ... but one: not black
    tile[:, 1, 0] = np.float32(-0.01)                        # EMIT masked reflectance
    tile[:, 1, 1] = np.float32(-0.0104)                      # within atol + rtol*|y| of -0.01
    tile[:, 1, 2] = np.float32(-0.0112)                      # just outside
    tile[:, 2, 0] = 0.0
    tile[:, 2, 1] = np.float32(9e-7)
    tile[:, 2, 2] = np.float32(1e-6)                         # not < float32(1e-6)?  decided by numpy
    tile[:, 3, 0] = np.float32(-9999.0009)                   # isclose to nodata in float32
    tile[:, 3, 1] = np.float32(-9999.2)
    tile[:, 4, 0] = np.nan
    tile[:, 4, 1] = -9999.0
    tile[2, 4, 1] = np.nan
    black_nd = ref.is_black_mask(tile, nodata=-9999.0)
    black_none = ref.is_black_mask(tile)
    black_custom = ref.is_black_mask(tile, nodata=0.5, masked_val=0.25, nodata_atol=0.3, zero_atol=0.05)
    picks = {f"idx_{n}_{k}": ref._subsample_bands_evenly(n, k) for n, k in ((285, 32), (285, 10), (40, 32), (32, 32), (33, 32))}
    np.savez_compressed(os.path.join(OUT, "tiles.npz"), tile=tile, black_nd=black_nd, black_none=black_none,
                        black_custom=black_custom, **picks)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()

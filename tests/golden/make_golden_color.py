"""Generate tests/golden/color_ot_match.npz by running the REFERENCE'S OWN ``ot_match_rgb_sinkhorn_pot``
(/root/reference/s2_emit/color.py:63-116) on small seeded inputs, its ``ot.dist`` / ``ot.sinkhorn`` calls resolved to
oracle/ot.py (POT is absent: parity unpinned for those two calls only).  Run in the build container; the reference is
not available on the GPU box.    python tests/golden/make_golden_color.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    ref = ref_loader.load()
    rng = np.random.default_rng(20261018)
    H, W = 48, 41
    src = (rng.random((H, W, 3)) ** 1.4 * np.array([0.9, 0.75, 0.6])).astype(np.float32)
    A = np.array([[0.9, 0.05, 0.0], [0.1, 0.8, 0.1], [0.0, 0.1, 0.7]])
    refimg = np.clip(src.astype(np.float64) @ A + np.array([0.03, 0.05, 0.08]) + rng.normal(0, 0.02, src.shape), 0, 1)
    refimg = refimg.astype(np.float32)
    src[3, 4, 2] = np.nan                  # dropped from X_all only; the pixel itself maps to NaN (it is in the mask)
    refimg[6, 7, 0] = np.inf               # dropped from Y_all only
    mask = rng.random((H, W)) < 0.75
    mask[3, 4] = mask[6, 7] = True
    out = {}
    with np.errstate(invalid="ignore"):
        out["out_n400_s0"] = ref.ot_match_rgb_sinkhorn_pot(src, refimg, mask, n_samples=400, seed=0)
        out["out_n100000_s2"] = ref.ot_match_rgb_sinkhorn_pot(src, refimg, mask, n_samples=100000, seed=2)   # all rows
        out["out_reg01_it20"] = ref.ot_match_rgb_sinkhorn_pot(src, refimg, mask, n_samples=300, reg=0.1, numItermax=20,
                                                              stopThr=0.0, seed=5)
    one = np.zeros((H, W), bool)
    one[10, 10] = True                     # fewer than 2 samples: a copy of the input (:86-87)
    out["out_one"] = ref.ot_match_rgb_sinkhorn_pot(src, refimg, one)
    np.savez_compressed(os.path.join(OUT, "color_ot_match.npz"), src=src, ref=refimg, mask=mask, one=one, **out)
    print("wrote color_ot_match.npz", {k: v.shape for k, v in out.items()})
    # ---- robust_norm / robust_norm_rgb (color.py:6-23): float64 stretches
    img = (rng.random((33, 29, 3)) ** 2 * np.array([0.7, 0.4, 0.25])).astype(np.float32)
    img[..., 2] = np.round(img[..., 2] * 40) / 40                  # ties
    rmask = rng.random((33, 29)) < 0.6
    xn = img[..., 0].copy()
    xn[rng.random(xn.shape) < 0.1] = np.nan                        # nanpercentile ignores them, the output keeps them
    xn[0, 0] = np.inf
    with np.errstate(invalid="ignore"):
        np.savez_compressed(os.path.join(OUT, "color_robust.npz"), img=img, mask=rmask, xn=xn,
                            rn=ref.robust_norm(xn), rn_5_90=ref.robust_norm(xn, 5, 90), rn_cube=ref.robust_norm(img),
                            rgb=ref.robust_norm_rgb(img, rmask), rgb_1_99=ref.robust_norm_rgb(img, rmask, 1, 99))
    print("wrote color_robust.npz")
    # ---- histogram matching (color.py:36-63)
    hs = (rng.random((40, 37, 3)) ** 1.7 * np.array([0.95, 0.8, 0.6])).astype(np.float32)
    hr = np.clip(0.1 + 0.8 * rng.random((40, 37, 3)) ** 0.6, 0, 1).astype(np.float32)
    hs[..., 1] = np.round(hs[..., 1] * 30) / 30                    # heavy ties in the source ...
    hr[..., 2] = np.round(hr[..., 2] * 12) / 12                    # ... and in the reference
    hmask = rng.random((40, 37)) < 0.7
    hs[~hmask] *= 1.8                                              # values above 1 outside the mask: clipped too
    few = np.zeros((40, 37), bool)
    few[5, 5] = few[6, 9] = True
    np.savez_compressed(os.path.join(OUT, "color_histmatch.npz"), src=hs, ref=hr, mask=hmask, few=few,
                        out=ref.histogram_match_rgb(hs, hr, hmask), out_few=ref.histogram_match_rgb(hs, hr, few))
    print("wrote color_histmatch.npz")


if __name__ == "__main__":
    main()

"""What an invalid (nodata) tile costs in the fused gather + SRF kernel: the granule's GLT, an all-nodata GLT of the same
shape, and the granule's GLT with its nodata corners cropped away (rows / columns of the bounding box only).
   python profiles/prof_invalid.py [reps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hsr_b200 import kernels, synthetic                     # noqa: E402
from hsr_b200.s2_emit.srf import srf_fold_weights, synthetic_s2_srf   # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
dev = torch.device("cuda:0")
Hr, Wr, B = synthetic.GRANULE_RAW_SHAPE
w = synthetic.emit_wavelengths()
good = synthetic.good_band_mask(w)
raw = synthetic.raw_cube_spectra_torch((Hr, Wr, B), 0, dev, good)
gx_np, gy_np = synthetic.rotation_glt(Hr, Wr, 25.0)
W, names, _, fo = srf_fold_weights(w, synthetic_s2_srf(), good)
Wt, fod = torch.from_numpy(W).to(dev), torch.from_numpy(fo).to(dev)
K = len(names)


def timed(gx, gy, label):
    Ho, Wo = gx.shape
    out = kernels.alloc_planes(K, (Ho, Wo), dev)
    fm = torch.empty((Ho, Wo), dtype=torch.bool, device=dev)
    f = lambda: kernels.glt_srf(raw, gx, gy, Wt, fod, bands_out=out, want_diag=False, fit_mask_out=fm, gate_k=0)
    for _ in range(5):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nv = int(((gx > 0) & (gy > 0)).sum())
    n = Ho * Wo
    print(f"{label:58s} {ms:7.4f} ms   {n:9d} px, {nv:9d} valid, {(n - nv) / 32 / 148:7.1f} nodata tiles per CTA", flush=True)
    return ms


gx = torch.from_numpy(gx_np).to(dev)
gy = torch.from_numpy(gy_np).to(dev)
t_all = timed(gx, gy, "granule GLT (25 deg, 43 % nodata)")
t_inv = timed(torch.zeros_like(gx), torch.zeros_like(gy), "all-nodata GLT, same shape")
# same valid pixels packed densely: the valid entries of every row moved to the row's left, rows then cut to the longest run
ok = (gx_np > 0) & (gy_np > 0)
cnt = ok.sum(1)
Wc = int(cnt.max())
gxc = np.zeros((gx_np.shape[0], Wc), np.int32)
gyc = np.zeros_like(gxc)
for r in range(gx_np.shape[0]):
    gxc[r, :cnt[r]] = gx_np[r, ok[r]]
    gyc[r, :cnt[r]] = gy_np[r, ok[r]]
t_c = timed(torch.from_numpy(gxc).to(dev), torch.from_numpy(gyc).to(dev), "same valid pixels, rows left-packed (fewer nodata tiles)")
print(f"nodata tiles: {t_inv / ((gx.numel()) / 32 / 148) * 1e6:.1f} ns per tile and CTA when the grid holds nothing else")

"""Decomposition of warp_pair_kernel with the EXPERIMENT library (HSR_B200_EXPERIMENTAL_LIB=1 HSR_WARP_DRY=bits):
1 no taps, 2 no stores, 4 no staging.  One process per setting (the knob is read per call, the library once).
    HSR_B200_EXPERIMENTAL_LIB=1 HSR_WARP_DRY=3 python profiles/prof_warp_dry.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hsr_b200 import kernels, synthetic  # noqa: E402
from hsr_b200.EMIT_data import warp as hwarp  # noqa: E402

Hr, Wr, B = 1280, 1242, 285
raw = synthetic.raw_cube_spectra_torch((Hr, Wr, B), 0, "cuda")
gx, gy = (torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in synthetic.rotation_glt(Hr, Wr, 25.0))
P = kernels.padded_bands(B)
buf = torch.empty((gx.shape[0], gx.shape[1], P), dtype=torch.float32, device="cuda")
_, valid, _ = kernels.glt_ortho(raw, gx, gy, out=buf, out_pix_stride=P)
Ho, Wo = valid.shape
ortho = buf[:, :, :B]
src_gt = (-118.60, 0.000542232520256367, 0.0, 34.90, 0.0, -0.000542232520256367)
s2 = hwarp.S2Grid(epsg=32611, x0=300000.0, y0=3900000.0, dx=10.0, dy=10.0, width=10980, height=10980)
dst_gt, (Hd, Wd), _ = hwarp.target_grid(src_gt, (Ho, Wo), s2)
scales = hwarp.warp_scales(dst_gt, src_gt, (Hd, Wd), 11, False)
out = torch.empty((Hd, Wd, P), dtype=torch.float32, device="cuda")[:, :, :B]
ws = torch.empty(Hd * Wd * 2, dtype=torch.float64, device="cuda")


def run():
    kernels.warp(ortho, src_gt, dst_gt, (Hd, Wd), utm_zone=11, nodata=-9999.0, out=out, kernel="cubic", scales=scales)


for _ in range(3):
    run()
torch.cuda.synchronize()
ts = []
for _ in range(8):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    run()
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print(f"HSR_WARP_DRY={os.environ.get('HSR_WARP_DRY', '0')}: {np.median(ts):.3f} ms (incl. 0.13 ms warp_coords)")

"""One cubic warp of the granule's ortho cube onto the UTM 60 m grid (ncu capture target for warp_pair_kernel).
    python profiles/prof_warp_once.py [kernel=cubic]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hsr_b200 import kernels, synthetic  # noqa: E402
from hsr_b200.EMIT_data import warp as hwarp  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "cubic"
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
for _ in range(2):
    kernels.warp(ortho, src_gt, dst_gt, (Hd, Wd), utm_zone=11, nodata=-9999.0, out=out, kernel=kind, scales=scales)
torch.cuda.synchronize()
print("ok")

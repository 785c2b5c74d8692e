"""Timing of the general grid warp (hsr_warp_f32) on the granule workload: the 1685 x 1667 x 285 WGS-84 ortho cube
(records padded to 288 floats) -> the snapped UTM 60 m grid, cubic, nodata -9999.  CUDA events, warm-up, medians.
    python profiles/prof_warp.py [reps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from hsr_b200 import kernels, synthetic
from hsr_b200.EMIT_data import warp as hwarp

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
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


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


print(f"ortho {Ho}x{Wo}x{B} -> utm {Hd}x{Wd}, scales {scales[0]:.4f} {scales[1]:.4f}")
nvalid = int(valid.sum().item())
for name, kw in (("cubic", dict(kernel="cubic", scales=scales)), ("cubic 4x4 (scale 1)", dict(kernel="cubic", scales=(1.0, 1.0))),
                 ("bilinear", dict(kernel="bilinear", scales=scales))):
    ms = timed(lambda: kernels.warp(ortho, src_gt, dst_gt, (Hd, Wd), utm_zone=11, nodata=-9999.0, out=out, **kw))
    covered = int((out[..., 0] != -9999.0).sum().item())
    # algorithmic bytes: every valid source spectrum read once + every destination spectrum written once
    gb = (nvalid * B * 4 + Hd * Wd * B * 4) / 1e9
    print(f"warp {name:22s} {ms:8.3f} ms  {gb / ms * 1e3:8.1f} GB/s algorithmic  ({Hd * Wd / ms / 1e3:.1f} Mpix/s, covered {covered / (Hd * Wd):.3f})")
unpadded = ortho.contiguous()
out2 = torch.empty((Hd, Wd, B), dtype=torch.float32, device="cuda")
ms = timed(lambda: kernels.warp(unpadded, src_gt, dst_gt, (Hd, Wd), utm_zone=11, nodata=-9999.0, scales=scales, out=out2))
print(f"warp cubic, scalar path (285-float records) {ms:8.3f} ms")
ms = timed(lambda: kernels.warp_coords(src_gt, dst_gt, (Hd, Wd), utm_zone=11))
print(f"warp_coords (transformer only)             {ms:8.3f} ms")

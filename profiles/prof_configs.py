"""Throughput of the other BASELINE configs on one GPU (parity-test cases, timed for the record):
config 3 (512 tiles 256 x 256 x 285, per-tile fits) and config 5 (8192^2 mosaic over a 6164^2 raw cube, fused gather + SRF).
    python profiles/prof_configs.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hsr_b200 import kernels, synthetic  # noqa: E402
from hsr_b200.pipeline import PairSynthesizer  # noqa: E402
from hsr_b200.s2_emit.srf import synthetic_s2_srf  # noqa: E402

dev = torch.device("cuda", 0)
w = synthetic.emit_wavelengths()
good = synthetic.good_band_mask(w)
ps = PairSynthesizer(w, synthetic_s2_srf(), good, deg=2, device=dev)
g = torch.Generator(device=dev).manual_seed(0)
B = 285


def timed(name, fn, npx, nbytes, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:64s} {ms:9.3f} ms   {npx / ms / 1e3:9.1f} Mpix/s   {nbytes / ms / 1e6:8.1f} GB/s algorithmic")


# ---- config 3: 512 tiles
T, h = 512, 256
raw = torch.empty((T, h, h, B), dtype=torch.float32, device=dev)
for t0 in range(0, T, 64):
    raw[t0:t0 + 64] = torch.rand((64, h, h, B), generator=g, device=dev) * 0.6
ii = torch.arange(h, device=dev, dtype=torch.int32)
gy = (ii.view(1, h, 1) + 1).expand(T, h, h).contiguous()
gx = (ii.view(1, 1, h) + 1).expand(T, h, h).contiguous()
gx[torch.rand((T, h, h), generator=g, device=dev) < 0.02] = 0
s2 = torch.rand((ps.K, T, h, h), generator=g, device=dev)
n = T * h * h
timed("config 3: 512 tiles 256x256x285, ortho+SRF+per-tile fit+apply", lambda: ps.synthesize_tiles(raw, gx, gy, s2), n,
      0.98 * n * B * 4 + n * 8 + n * ps.K * 4 * 5 + 4 * n)
del raw, gx, gy, s2
torch.cuda.empty_cache()

# ---- config 5: mosaic
Ho = Wo = 8192
Hr = Wr = 6164
raw = torch.empty((Hr, Wr, B), dtype=torch.float32, device=dev)
for r0 in range(0, Hr, 512):
    r1 = min(Hr, r0 + 512)
    raw[r0:r1] = torch.rand((r1 - r0, Wr, B), generator=g, device=dev) * 0.6
th = np.deg2rad(25.0)
yy = torch.arange(Ho, device=dev, dtype=torch.float64).view(-1, 1) - (Ho - 1) / 2
xx = torch.arange(Wo, device=dev, dtype=torch.float64).view(1, -1) - (Wo - 1) / 2
rx = torch.round(xx * np.cos(th) + yy * np.sin(th) + (Wr - 1) / 2).to(torch.int64)
ry = torch.round(-xx * np.sin(th) + yy * np.cos(th) + (Hr - 1) / 2).to(torch.int64)
inside = (rx >= 0) & (rx < Wr) & (ry >= 0) & (ry < Hr)
gx = torch.where(inside, rx + 1, torch.zeros_like(rx)).to(torch.int32)
gy = torch.where(inside, ry + 1, torch.zeros_like(ry)).to(torch.int32)
n, nv = Ho * Wo, int(inside.sum())
del rx, ry, xx, yy, inside
bands = kernels.alloc_planes(ps.K, (Ho, Wo), dev)
fm = torch.empty((Ho, Wo), dtype=torch.bool, device=dev)
timed("config 5: 8192^2 mosaic (43 GB raw), fused gather + SRF + fit mask",
      lambda: ps.bands_from_raw(raw, gx, gy, bands_out=bands, fit_mask_out=fm), n, nv * B * 4 + n * 8 + n * ps.K * 4 + 2 * n, reps=3)
print("peak HBM in use: %.1f GB" % (torch.cuda.max_memory_allocated() / 2 ** 30))

"""Timing of the "next row" kernels at granule scale: masked percentiles + stretch (color.py:25-34), tile
validity / quantisation (tiles_helpers/utils.py), OT targets.   python profiles/prof_next.py [reps]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hsr_b200 import kernels  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda", 0)
Ho, Wo = 1685, 1667
n = Ho * Wo
g = torch.Generator(device=dev).manual_seed(0)


def timed(name, fn, nbytes):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:58s} {ms:8.3f} ms   {nbytes / ms / 1e6:8.1f} GB/s algorithmic")


mask = torch.rand((Ho, Wo), generator=g, device=dev) < 0.566
for K in (3, 12):
    x = kernels.alloc_planes(K, (Ho, Wo), dev)
    x.copy_(torch.rand((K, Ho, Wo), generator=g, device=dev) ** 2)
    # one streaming pass over the planes (+ the mask once per series) since round 2; bytes actually read
    timed(f"masked_percentiles [2, 98] (K = {K}, one pass)", lambda: kernels.masked_percentiles(x, mask, [2, 98]),
          n * K * 4 + n * K)
    lohi = kernels.masked_percentiles(x, mask, [2, 98])
    out = kernels.alloc_planes(K, (Ho, Wo), dev)
    timed(f"stretch_apply (K = {K})", lambda: kernels.stretch_apply(x, lohi, out=out), 2 * n * K * 4)
    del x, out
# tile validity / quantisation on a band-sequential granule-sized cube (285 bands)
B = 285
cube = torch.rand((B, Ho, Wo), generator=g, device=dev) * 0.6
cube[:, :400, :] = -9999.0
# bytes actually fetched: the 400 all-nodata rows stay black to the last band (all B bands are read), a random pixel
# stops being black at its first band (one 32-byte sector per 8 pixels of band 0 and of nothing else)
n_fill = 400 * Wo
timed("black_mask (285 bands; nodata rows read all bands, the rest exits after 1)", lambda: kernels.black_mask(cube, -9999.0),
      n_fill * B * 4 + (n - n_fill) * 4 + n)
timed("quantize_u16 (285 bands)", lambda: kernels.quantize_u16(cube, -9999.0), n * B * 6)
bm = kernels.black_mask(cube, -9999.0)
timed("tile_sums (100 x 100 windows)", lambda: kernels.tile_sums(bm, 100, 100), n)

# fused tile export from the raw granule (BIP) through the GLT: gather + quantise + black mask in one pass
from hsr_b200 import synthetic  # noqa: E402
Hr, Wr, Bb = synthetic.GRANULE_RAW_SHAPE
del cube, bm
raw = torch.rand((Hr, Wr, Bb), generator=g, device=dev) * 0.6
gx_np, gy_np = synthetic.rotation_glt(Hr, Wr, 25.0)
gx, gy = torch.from_numpy(gx_np).to(dev), torch.from_numpy(gy_np).to(dev)
n_v = int(((gx_np != 0) & (gy_np != 0)).sum())
timed("glt_ortho_u16 (fused gather + u16 + black mask)", lambda: kernels.glt_ortho_u16(raw, gx, gy, want_diag=False),
      n_v * Bb * 4 + n * Bb * 2 + n * 8 + 2 * n)
timed("glt_ortho_u16 (no black mask)", lambda: kernels.glt_ortho_u16(raw, gx, gy, want_diag=False, want_black=False),
      n_v * Bb * 4 + n * Bb * 2 + n * 8 + n)

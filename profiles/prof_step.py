"""One pass of the benchmark step (4 launches: glt_stream<SRF>, poly_moments, moments_finalize, solve_apply)
at full granule size — the ncu capture target for the whole hot path.
    python profiles/prof_step.py [reps]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hsr_b200 import kernels, synthetic  # noqa: E402
from hsr_b200.pipeline import PairSynthesizer  # noqa: E402
from hsr_b200.s2_emit.srf import synthetic_s2_srf  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda", 0)
Hr, Wr, B = synthetic.GRANULE_RAW_SHAPE
w = synthetic.emit_wavelengths()
good = synthetic.good_band_mask(w)
ps = PairSynthesizer(w, synthetic_s2_srf(), good, deg=2, device=dev)
raw = synthetic.raw_cube_spectra_torch((Hr, Wr, B), 0, dev, good)
gx_np, gy_np = synthetic.rotation_glt(Hr, Wr, 25.0)
gx, gy = torch.from_numpy(gx_np).to(dev), torch.from_numpy(gy_np).to(dev)
Ho, Wo = gx_np.shape
bands = kernels.alloc_planes(ps.K, (Ho, Wo), dev)
matched = kernels.alloc_planes(ps.K, (Ho, Wo), dev)
b, valid, _, _ = ps.bands_from_raw(raw, gx, gy, bands_out=bands)
s2 = kernels.alloc_planes(ps.K, (Ho, Wo), dev)
s2.copy_(synthetic.s2_reference_torch(b, seed=1))
fmask = torch.empty((Ho, Wo), dtype=torch.bool, device=dev)
torch.cuda.synchronize()
lo, hi = ps.clip
for _ in range(reps):
    b, valid, _, _ = ps.bands_from_raw(raw, gx, gy, bands_out=bands, fit_mask_out=fmask)
    mom, fm, _, _ = ps.fit(b, s2, valid, fmask)
    kernels.poly_solve_apply(b, mom, fm, 2, min_count=ps.min_count, lo=lo, hi=hi, out=matched)
torch.cuda.synchronize()
print("ok")

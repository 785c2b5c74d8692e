"""A few steps of the stretch variant of the granule pass (ncu launch-list target): SRF -> percentiles of both images
-> moments on stretched values -> solve + apply on stretched values.
    python profiles/prof_stretch_step.py [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hsr_b200 import kernels, synthetic  # noqa: E402
from hsr_b200.pipeline import PairSynthesizer  # noqa: E402
from hsr_b200.s2_emit.srf import synthetic_s2_srf  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda", 0)
Hr, Wr, B = synthetic.GRANULE_RAW_SHAPE
w = synthetic.emit_wavelengths()
good = synthetic.good_band_mask(w)
ps = PairSynthesizer(w, synthetic_s2_srf(), good, deg=2, device=dev, stretch=(2, 98))
raw = synthetic.raw_cube_spectra_torch((Hr, Wr, B), 0, dev, good)
gx_np, gy_np = synthetic.rotation_glt(Hr, Wr, 25.0)
gx, gy = torch.from_numpy(gx_np).to(dev), torch.from_numpy(gy_np).to(dev)
b0 = ps.bands_from_raw(raw, gx, gy)[0]
s2 = kernels.alloc_planes(ps.K, gx.shape, dev)
s2.copy_(synthetic.s2_reference_torch(b0, seed=1))
del b0
for _ in range(steps):
    res = ps.synthesize(raw, gx, gy, s2)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    res = ps.synthesize(raw, gx, gy, s2)
e1.record()
torch.cuda.synchronize()
print(f"stretch step (eager): {e0.elapsed_time(e1) / 10:.4f} ms; limits band0 {res.x_limits[0].tolist()}")

"""glt_ortho_u16 (fused tile export) on the granule, a few launches — for an ncu capture of glt_stream_kernel<Q16>.
   python profiles/prof_q16_once.py [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hsr_b200 import kernels, synthetic                     # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda:0")
Hr, Wr, B = synthetic.GRANULE_RAW_SHAPE
w = synthetic.emit_wavelengths()
raw = synthetic.raw_cube_spectra_torch((Hr, Wr, B), 0, dev, synthetic.good_band_mask(w))
gx_np, gy_np = synthetic.rotation_glt(Hr, Wr, 25.0)
gx, gy = torch.from_numpy(gx_np).to(dev), torch.from_numpy(gy_np).to(dev)
for variant in (dict(), dict(want_black=False)):
    kernels.glt_ortho_u16(raw, gx, gy, want_diag=False, **variant)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        kernels.glt_ortho_u16(raw, gx, gy, want_diag=False, **variant)
    e1.record()
    torch.cuda.synchronize()
    print(variant, f"{e0.elapsed_time(e1) / reps:.4f} ms", flush=True)

"""One launch of each glt_stream flavour at full granule size (ncu capture target):
fused glt_srf, glt_ortho (materialise), un-fused srf on the ortho cube."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hsr_b200 import kernels, synthetic  # noqa: E402
from hsr_b200.s2_emit.srf import srf_fold_weights, synthetic_s2_srf  # noqa: E402

dev = torch.device("cuda", 0)
Hr, Wr, B = synthetic.GRANULE_RAW_SHAPE
w = synthetic.emit_wavelengths()
good = synthetic.good_band_mask(w)
raw = synthetic.raw_cube_spectra_torch((Hr, Wr, B), 0, dev, good)
gx_np, gy_np = synthetic.rotation_glt(Hr, Wr, 25.0)
gx, gy = torch.from_numpy(gx_np).to(dev), torch.from_numpy(gy_np).to(dev)
W, names, _, fo = srf_fold_weights(w, synthetic_s2_srf(), good)
Wd, fod = torch.from_numpy(W).to(dev), torch.from_numpy(fo).to(dev)
Ho, Wo = gx_np.shape
bands = torch.empty((len(names), Ho, Wo), dtype=torch.float32, device=dev)
out = torch.empty((Ho, Wo, B), dtype=torch.float32, device=dev)
torch.cuda.synchronize()
kernels.glt_srf(raw, gx, gy, Wd, fod, bands_out=bands, want_diag=False)
kernels.glt_ortho(raw, gx, gy, out=out, want_diag=False)
kernels.srf_integrate(out, Wd, bands_out=bands)
torch.cuda.synchronize()
print("ok")

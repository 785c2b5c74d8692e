"""Run each hot-path kernel a few times at full granule size (for ncu captures and quick timing).
    python profiles/prof_kernels.py [srf|ortho|poly|all] [reps]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from hsr_b200 import kernels, synthetic  # noqa: E402
from hsr_b200.s2_emit.srf import srf_fold_weights, synthetic_s2_srf  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
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
n_o, n_v, K = Ho * Wo, int(((gx_np != 0) & (gy_np != 0)).sum()), len(names)


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
    print(f"{name:36s} {ms:8.3f} ms   {nbytes / ms / 1e6:8.1f} GB/s algorithmic")


bands = kernels.alloc_planes(K, (Ho, Wo), dev)
if which in ("srf", "all"):
    timed("glt_srf (fused)", lambda: kernels.glt_srf(raw, gx, gy, Wd, fod, bands_out=bands, want_diag=False),
          n_v * B * 4 + n_o * 8 + n_o * K * 4 + n_o)
    fmo = torch.empty((Ho, Wo), dtype=torch.bool, device=dev)
    timed("glt_srf (fused, + fit mask)", lambda: kernels.glt_srf(raw, gx, gy, Wd, fod, bands_out=bands, want_diag=False,
                                                                 fit_mask_out=fmo, gate_k=0),
          n_v * B * 4 + n_o * 8 + n_o * K * 4 + 2 * n_o)
if which in ("ortho", "all"):
    out = torch.empty((Ho, Wo, B), dtype=torch.float32, device=dev)
    timed("glt_ortho (materialise)", lambda: kernels.glt_ortho(raw, gx, gy, out=out, want_diag=False),
          n_v * B * 4 + n_o * B * 4 + n_o * 8 + n_o)
    timed("srf (un-fused, ortho cube)", lambda: kernels.srf_integrate(out, Wd, bands_out=bands),
          n_o * B * 4 + n_o * K * 4)
    del out
if which in ("poly", "all"):
    kernels.glt_srf(raw, gx, gy, Wd, fod, bands_out=bands)
    s2 = kernels.alloc_planes(K, (Ho, Wo), dev)
    s2.copy_(synthetic.s2_reference_torch(bands, seed=1))
    valid = ((gx != 0) & (gy != 0))
    fm = kernels.fit_mask(bands, valid)
    timed("fit_mask", lambda: kernels.fit_mask(bands, valid), n_o * K * 4 + 2 * n_o)
    timed("poly_moments (deg 2)", lambda: kernels.poly_moments(bands, s2, fm, 2), 2 * n_o * K * 4 + n_o)
    mom = kernels.poly_moments(bands, s2, fm, 2)
    timed("poly_solve", lambda: kernels.poly_solve(mom, 2, 200), K * 8 * 11)
    co = kernels.poly_solve(mom, 2, 200)
    o2 = kernels.alloc_planes(K, (Ho, Wo), dev)
    timed("poly_apply", lambda: kernels.poly_apply(bands, co, fm, out=o2), 2 * n_o * K * 4 + n_o)
    timed("fit_moments (mask + moments)", lambda: kernels.fit_moments(bands, s2, valid, 2),
          3 * n_o * K * 4 + 3 * n_o)
    timed("fit_moments (mask given)", lambda: kernels.fit_moments(bands, s2, fm, 2, mask_given=True),
          2 * n_o * K * 4 + n_o)
    mom2, fm2 = kernels.fit_moments(bands, s2, valid, 2)
    timed("poly_solve_apply (fused)", lambda: kernels.poly_solve_apply(bands, mom2, fm2, 2, min_count=200, out=o2),
          2 * n_o * K * 4 + n_o)

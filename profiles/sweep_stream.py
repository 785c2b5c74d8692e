"""Sweep the launch knobs of glt_stream_kernel (producer warps, stages, run merging) at full granule
size.  The knobs are read from the environment by the library at every launch (HSR_PRODUCERS,
HSR_STAGES, HSR_NO_MERGE); defaults are what ships.
    python profiles/sweep_stream.py [theta_deg ...]
"""
import itertools
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from hsr_b200 import kernels, synthetic  # noqa: E402
from hsr_b200.s2_emit.srf import srf_fold_weights, synthetic_s2_srf  # noqa: E402

thetas = [float(a) for a in sys.argv[1:]] or [25.0]
dev = torch.device("cuda", 0)
Hr, Wr, B = synthetic.GRANULE_RAW_SHAPE
w = synthetic.emit_wavelengths()
good = synthetic.good_band_mask(w)
raw = synthetic.raw_cube_spectra_torch((Hr, Wr, B), 0, dev, good)
W, names, _, fo = srf_fold_weights(w, synthetic_s2_srf(), good)
Wd, fod = torch.from_numpy(W).to(dev), torch.from_numpy(fo).to(dev)
K = len(names)


def timed(fn, reps=10):
    fn()
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for theta in thetas:
    gx_np, gy_np = synthetic.rotation_glt(Hr, Wr, theta)
    gx, gy = torch.from_numpy(gx_np).to(dev), torch.from_numpy(gy_np).to(dev)
    Ho, Wo = gx_np.shape
    n_o, n_v = Ho * Wo, int(((gx_np != 0) & (gy_np != 0)).sum())
    bands = torch.empty((K, Ho, Wo), dtype=torch.float32, device=dev)
    out = torch.empty((Ho, Wo, B), dtype=torch.float32, device=dev)
    srf_bytes = n_v * B * 4 + n_o * 8 + n_o * K * 4 + n_o
    cp_bytes = n_v * B * 4 + n_o * B * 4 + n_o * 8 + n_o
    print(f"theta {theta}: ortho {Ho}x{Wo}, valid {n_v / n_o:.3f}")
    for nomerge, nprod, nst in itertools.product((0, 1), (3, 4, 5), (5, 4)):
        os.environ["HSR_NO_MERGE"] = str(nomerge)
        os.environ["HSR_PRODUCERS"] = str(nprod)
        os.environ["HSR_STAGES"] = str(nst)
        t_srf = timed(lambda: kernels.glt_srf(raw, gx, gy, Wd, fod, bands_out=bands, want_diag=False))
        t_cp = timed(lambda: kernels.glt_ortho(raw, gx, gy, out=out, want_diag=False), reps=5) if nst == 5 else float("nan")
        print(f"  merge={1 - nomerge} prod={nprod} stages={nst}:  glt_srf {t_srf:7.3f} ms {srf_bytes / t_srf / 1e6:7.1f} GB/s"
              f"   glt_ortho {t_cp:7.3f} ms {cp_bytes / t_cp / 1e6:7.1f} GB/s", flush=True)
    for k in ("HSR_NO_MERGE", "HSR_PRODUCERS", "HSR_STAGES"):
        os.environ.pop(k, None)
    t_un = timed(lambda: kernels.srf_integrate(out, Wd, bands_out=bands))
    print(f"  srf (un-fused, identity runs) {t_un:7.3f} ms {(n_o * B * 4 + n_o * K * 4) / t_un / 1e6:7.1f} GB/s")

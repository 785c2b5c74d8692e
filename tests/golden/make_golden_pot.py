"""Pin the Sinkhorn boundary the moment POT is importable:

    python tests/golden/make_golden_pot.py      ->  tests/golden/pot_sinkhorn.npz

Runs the REAL third-party package the reference imports (`import ot`, s2_emit/poly_regression.py:4,52-53;
s2_emit/color.py:3,100-101) on seeded inputs and stores inputs + outputs + the POT version.  POT is not installed in the
build image and there is no wheel in /opt/wheelhouse, so until someone runs this where `pip install pot` is possible the
file is absent, tests/test_pot_pin.py SKIPS with "parity unpinned", and oracle/ot.py (a restatement of POT's published
algorithm) stays the only checker — cross-checked meanwhile by an independent long-double log-domain solve
(tests/test_oracle_golden.py::test_sinkhorn_fixed_point_against_independent_logdomain_solve).
When /root/reference is also present the reference's own fit_ot_poly_rgb / ot_match_rgb_sinkhorn_pot are run with the
real POT and stored as well."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pot_sinkhorn.npz")


def main():
    try:
        import ot
    except ImportError:
        print("POT (import name `ot`) is not installed: nothing generated; parity for ot.dist / ot.sinkhorn stays unpinned")
        return 1
    rng = np.random.default_rng(20240820)
    save = {"pot_version": np.array(getattr(ot, "__version__", "unknown"))}
    cases = {"small": (64, 60), "mid": (500, 430), "ragged": (257, 1001)}
    for name, (ns, nt) in cases.items():
        X = rng.random((ns, 3))
        Y = rng.random((nt, 3)) ** 1.5 * 0.9 + 0.05
        a = np.full(ns, 1.0 / ns)
        b = np.full(nt, 1.0 / nt)
        M = ot.dist(X, Y, metric="sqeuclidean")                                   # poly_regression.py:52
        P = ot.sinkhorn(a, b, M, reg=0.05, numItermax=300, stopThr=1e-6)          # :53
        P12 = ot.sinkhorn(a, b, M, reg=0.05, numItermax=12, stopThr=0.0)
        save.update({f"{name}_X": X, f"{name}_Y": Y, f"{name}_M": np.asarray(M), f"{name}_P": np.asarray(P),
                     f"{name}_P12": np.asarray(P12),
                     f"{name}_ybar": (np.asarray(P) @ Y) / (np.asarray(P).sum(1, keepdims=True) + 1e-32)})
    if os.path.isdir("/root/reference"):
        from oracle import ref_loader

        ref = ref_loader.load(ot_module=ot)
        g = np.load(os.path.join(os.path.dirname(OUT), "ot_fit.npz"))
        save.update({"fit_src": g["src"], "fit_ref": g["ref"], "fit_mask": g["mask"]})
        for deg, n, seed in ((2, 400, 0), (4, 300, 1)):
            save[f"fit_coeffs_d{deg}_n{n}_s{seed}"] = ref.fit_ot_poly_rgb(g["src"], g["ref"], g["mask"], deg=deg,
                                                                            n_samples=n, seed=seed)
    np.savez_compressed(OUT, **save)
    print("wrote", OUT, "POT", save["pot_version"])
    return 0


if __name__ == "__main__":
    sys.exit(main())

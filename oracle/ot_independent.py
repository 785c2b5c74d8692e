"""An independent checker for oracle/ot.py (test infrastructure — see oracle/__init__.py).

oracle/ot.py restates POT's published scaling-form Sinkhorn-Knopp and stays "parity unpinned" until the real package can
produce golden vectors (tests/golden/make_golden_pot.py).  Meanwhile this module solves the SAME entropic optimal
transport problem by a different route, sharing no code and no formulation with it; since the optimum is unique, any
correct implementation run to convergence must land on the same plan.
"""
import numpy as np


def logdomain_sinkhorn_longdouble(X, Y, reg, iters):
    """An INDEPENDENT solve of the same entropic OT problem: log-domain (dual potentials f, g; logsumexp), numpy
    long double, direct squared distances (no |x|^2 + |y|^2 - 2xy expansion), run to its fixed point.  Shares no
    code and no formulation with oracle/ot.py (scaling vectors u, v on K = exp(-M / reg))."""
    ld = np.longdouble
    X, Y = X.astype(ld), Y.astype(ld)
    M = ((X[:, None, :] - Y[None, :, :]) ** 2).sum(-1)
    ns, nt = M.shape
    la, lb = np.log(ld(1) / ns), np.log(ld(1) / nt)
    f, g = np.zeros(ns, ld), np.zeros(nt, ld)

    def lse(A, axis):
        m = A.max(axis=axis, keepdims=True)
        return (m + np.log(np.exp(A - m).sum(axis=axis, keepdims=True))).squeeze(axis)

    for _ in range(iters):
        g = reg * (lb - lse((f[:, None] - M) / reg, 0))
        f = reg * (la - lse((g[None, :] - M) / reg, 1))
    P = np.exp((f[:, None] + g[None, :] - M) / reg)
    return P, M

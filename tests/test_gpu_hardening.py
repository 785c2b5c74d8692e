"""Round-2 hardening of the shipped library, on the GPU: the bounded peer wait (VERDICT r1 / ADVICE r1), the raw-NCCL
entry point hsr_allreduce_moments, host-side contract checks of the pair synthesizer."""
import ctypes

import numpy as np
import pytest
import torch

from hsr_b200 import _lib, kernels, synthetic
from hsr_b200 import dist as hdist
from hsr_b200.pipeline import PairSynthesizer
from hsr_b200.s2_emit.srf import synthetic_s2_srf

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _planes(K, n, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    x = kernels.alloc_planes(K, (n,), DEV)
    x.copy_(torch.rand((K, n), generator=g, device=DEV) * 0.7 + 0.05)
    y = kernels.alloc_planes(K, (n,), DEV)
    y.copy_(0.9 * x - 0.2 * x * x + 0.03)
    return x, y


def test_peer_wait_is_bounded_and_reports_the_missing_rank():
    """A world of two in which rank 1 never publishes: rank 0's solve/apply must give up after timeout_ms, write NaN
    coefficients (never a silent per-rank fit) and leave HSR_PEER_TIMEOUT in the block's status word; with every
    flag raised the same descriptor shape works.  One process, one GPU: both 'peer blocks' are this rank's own, so
    nothing waits on another launch (B200_PROFILING.md forbids kernels that wait on one another on one GPU)."""
    lib = _lib.lib()
    K, n, deg = 3, 50_000, 2
    x, y = _planes(K, n, 0)
    mask = torch.ones(n, dtype=torch.bool, device=DEV)
    blk = ctypes.c_void_p()
    _lib.check(lib.hsr_peer_alloc(ctypes.byref(blk)))
    try:
        peers = torch.tensor([blk.value, blk.value], dtype=torch.int64, device=DEV)
        word = ctypes.c_uint(7)
        _lib.check(lib.hsr_peer_status(blk.value, ctypes.byref(word), torch.cuda.current_stream().cuda_stream))
        assert word.value == 0
        ex = _lib.Exchange(peers.data_ptr(), blk.value, 2, 0, 0, 200, 0)       # 200 ms
        ex._stage = 0
        mom, _ = kernels.fit_moments(x, y, mask, deg, mask_given=True, exchange=ex)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        coeffs, out = kernels.poly_solve_apply(x, mom, mask, deg, exchange=ex)
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1)
        assert 150.0 < ms < 5000.0, f"gave up after {ms:.0f} ms, expected ~200"
        assert torch.isnan(coeffs).all(), "a fit without every rank's moments must not look like a result"
        rc = lib.hsr_peer_status(blk.value, ctypes.byref(word), torch.cuda.current_stream().cuda_stream)
        assert rc == _lib.HSR_EPEER and word.value & _lib.HSR_PEER_TIMEOUT
        assert b"timed out" in lib.hsr_last_error()
        with pytest.raises(_lib.HsrError):
            _lib.check(rc)
    finally:
        lib.hsr_peer_free(blk.value)
    # the same world with both ranks present, emulated in stream order on one GPU (host-numbered epoch 1): "rank 1"
    # publishes its sums, then rank 0 publishes and consumes — the system solved is the rank-ordered sum
    x1, y1 = _planes(K, n, 1)
    _lib.check(lib.hsr_peer_alloc(ctypes.byref(blk)))
    try:
        peers = torch.tensor([blk.value, blk.value], dtype=torch.int64, device=DEV)
        ex1 = _lib.Exchange(peers.data_ptr(), blk.value, 2, 1, 1, 2000, 0)
        ex1._stage = 0
        mom1, _ = kernels.fit_moments(x1, y1, mask, deg, mask_given=True, exchange=ex1)
        ex0 = _lib.Exchange(peers.data_ptr(), blk.value, 2, 0, 1, 2000, 0)
        ex0._stage = 0
        mom0, _ = kernels.fit_moments(x, y, mask, deg, mask_given=True, exchange=ex0)
        gm = torch.empty_like(mom0)
        coeffs, out = kernels.poly_solve_apply(x, mom0, mask, deg, exchange=ex0, moments_out=gm)
        torch.cuda.synchronize()
        assert torch.equal(gm, mom0 + mom1)                                    # slots added in rank order
        ref = kernels.poly_solve((mom0 + mom1).view(K, -1), deg)
        assert torch.equal(coeffs.view(K, -1), ref) and torch.isfinite(coeffs).all()
        _lib.check(lib.hsr_peer_status(blk.value, ctypes.byref(word), torch.cuda.current_stream().cuda_stream))
        assert word.value == 0
    finally:
        lib.hsr_peer_free(blk.value)


def test_allreduce_moments_through_a_raw_nccl_communicator():
    """hsr_allreduce_moments: ncclAllReduce(double, sum) on a communicator the host made without
    torch.distributed.  A one-rank communicator on one GPU: the sum is the input, the call goes through NCCL."""
    try:
        uid = hdist.RawNcclComm.unique_id()
    except RuntimeError as e:                                              # pragma: no cover
        pytest.skip(str(e))
    comm = hdist.RawNcclComm(uid, 1, 0)
    try:
        mom = torch.arange(96, dtype=torch.float64, device=DEV).view(12, 8) * 0.25
        want = mom.clone()
        comm.allreduce_moments(mom)
        torch.cuda.synchronize()
        assert torch.equal(mom, want)
        lib = _lib.lib()
        rc = lib.hsr_allreduce_moments(mom.data_ptr(), 96, None, torch.cuda.current_stream().cuda_stream)
        assert rc == -1 and b"null" in lib.hsr_last_error()
        with pytest.raises(TypeError):
            comm.allreduce_moments(mom.float())
    finally:
        comm.close()


def test_pair_synthesizer_contract_errors():
    w = synthetic.emit_wavelengths()
    good = synthetic.good_band_mask(w)
    with pytest.raises(KeyError, match="gate_band"):
        PairSynthesizer(w, synthetic_s2_srf(), good, gate_band="B99", device=DEV)
    ps = PairSynthesizer(w, synthetic_s2_srf(), good, gate_band="B4", device=DEV)
    assert ps.band_names[ps.gate_k] == "B4"
    ps = PairSynthesizer(w, synthetic_s2_srf(), good, stretch=(2, 98), device=DEV)
    with pytest.raises(ValueError, match="global fit"):
        ps.synthesize_sharded([])


def test_fit_returns_its_own_stretch_limits():
    """the percentile limits belong to the call, not to the instance."""
    w = synthetic.emit_wavelengths()
    good = synthetic.good_band_mask(w)
    ps = PairSynthesizer(w, synthetic_s2_srf(), good, stretch=(2, 98), device=DEV)
    Hr, Wr = 64, 48
    gx, gy = synthetic.rotation_glt(Hr, Wr, 25.0)
    gx, gy = torch.from_numpy(gx).to(DEV), torch.from_numpy(gy).to(DEV)
    res = []
    for seed in (0, 1):
        raw = synthetic.raw_cube_spectra_torch((Hr, Wr, 285), seed=seed, device=DEV, good=good)
        b0 = ps.bands_from_raw(raw, gx, gy)[0]
        s2 = synthetic.s2_reference_torch(b0, seed=seed)
        res.append(ps.synthesize(raw, gx, gy, s2))
    assert res[0].x_limits is not None and not torch.equal(res[0].x_limits, res[1].x_limits)
    assert not hasattr(ps, "_xl")
    for r in res:
        want = kernels.masked_percentiles(r.bands, r.fit_mask, (2, 98))
        assert torch.equal(r.x_limits, want)


def test_cuda_sinkhorn_against_independent_logdomain_solve():
    """the CUDA Sinkhorn against an INDEPENDENT formulation of the same problem (log-domain, long
    double, numpy; oracle/ot_independent.py) — not against oracle/ot.py.  Run to the fixed point, the barycentric
    targets of both must agree (the entropic optimum is unique)."""
    from oracle.ot_independent import logdomain_sinkhorn_longdouble as _logdomain_sinkhorn_longdouble

    rng = np.random.default_rng(42)
    for ns, nt in ((64, 64), (48, 82)):
        X = rng.random((ns, 3))
        Y = rng.random((nt, 3)) ** 1.3 * 0.8 + 0.1
        Pref, _ = _logdomain_sinkhorn_longdouble(X, Y, 0.05, 4000)
        want = ((Pref @ Y.astype(np.longdouble)) / Pref.sum(1, keepdims=True)).astype(np.float64)
        ybar, info = kernels.sinkhorn_barycentric(torch.from_numpy(X).to(DEV), torch.from_numpy(Y).to(DEV), 0.05, 20000, 1e-15)
        assert info.cpu().tolist()[3] == 0
        np.testing.assert_allclose(ybar.cpu().numpy(), want, rtol=0, atol=1e-12)

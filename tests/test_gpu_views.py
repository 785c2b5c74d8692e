"""hsr_raw_view_t on the GPU: raw-row windows (per-slab staging of a mosaic, SURVEY 7.3-6 / BASELINE configs[4]) and
in-kernel tile batches (configs[2]) must not change a single bit of the gather, the masks or the diagnostics."""
import numpy as np
import pytest
import torch

from hsr_b200 import dist as hdist
from hsr_b200 import kernels, synthetic
from hsr_b200.pipeline import PairSynthesizer
from hsr_b200.s2_emit import srf
from oracle import glt as oglt

pytestmark = pytest.mark.gpu
DEV = "cuda"


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _case(Hr, Wr, B, seed, theta=25.0):
    raw = synthetic.raw_cube_bits_np((Hr, Wr, B), seed=seed)
    gx, gy = synthetic.rotation_glt(Hr, Wr, theta)
    gx, gy = synthetic.inject_glt_defects(gx, gy, Hr, Wr, seed=seed + 1, n_oob=8, n_neg=8)
    return raw, gx, gy


def _np_range(gx, gy, Hr, Wr, transpose=False):
    ok = (gx != 0) & (gy != 0) & (gx - 1 >= 0) & (gx - 1 < Wr) & (gy - 1 >= 0) & (gy - 1 < Hr)
    s = (gx if transpose else gy)[ok] - 1
    return (int(s.min()), int(s.max()) + 1) if s.size else None


@pytest.mark.parametrize("transposed", [False, True])
@pytest.mark.parametrize("B", [285, 3])
def test_row_window_is_bit_identical_to_the_whole_cube(transposed, B):
    Hr, Wr = 90, 70
    raw, gx, gy = _case(Hr, Wr, B, seed=5)
    Ho, Wo = gx.shape
    mem = np.ascontiguousarray(np.transpose(raw, (1, 0, 2))) if transposed else raw
    full, vfull, dfull = kernels.glt_ortho(dev(mem), dev(gx), dev(gy), transpose_raw_yx=transposed)
    want, vref, _ = oglt.glt_ortho(raw, gx, gy)
    assert np.array_equal(full.cpu().numpy().view(np.int32), want.view(np.int32))
    for r0, r1 in [(0, Ho), (0, Ho // 3), (Ho // 3, 2 * Ho // 3), (2 * Ho // 3, Ho), (Ho // 2, Ho // 2 + 1)]:
        sx, sy = gx[r0:r1], gy[r0:r1]
        rng = kernels.glt_row_range(dev(sx), dev(sy), Hr, Wr, transpose_raw_yx=transposed).tolist()
        ref = _np_range(sx, sy, Hr, Wr, transposed)
        assert ref is not None and tuple(rng) == ref
        lo, hi = rng
        total = Wr if transposed else Hr
        out, v, d = kernels.glt_ortho(dev(mem[lo:hi]), dev(sx), dev(sy), transpose_raw_yx=transposed, raw_row0=lo,
                                      raw_rows_total=total)
        assert torch.equal(out.view(torch.int32), full[r0:r1].view(torch.int32)) and torch.equal(v, vfull[r0:r1])
        dd = d.tolist()
        assert dd[3] == 0 and dd[:3] == [int(((sx != 0) & (sy != 0)).sum()), int(vref[r0:r1].sum()),
                                           int(((sx != 0) & (sy != 0)).sum()) - int(vref[r0:r1].sum())]
        if hi - lo > 2:      # a window one row short at each end: those entries get the fill value and are COUNTED
            out2, v2, d2 = kernels.glt_ortho(dev(mem[lo + 1:hi - 1]), dev(sx), dev(sy), transpose_raw_yx=transposed,
                                             raw_row0=lo + 1, raw_rows_total=total)
            slow = (sx if transposed else sy) - 1
            outside = vref[r0:r1] & ((slow < lo + 1) | (slow >= hi - 1))
            assert int(d2[3]) == int(outside.sum()) > 0 and torch.equal(v2, vfull[r0:r1])
            got = out2.cpu().numpy()
            assert (got[outside] == -9999.0).all()
            keep = ~outside
            assert np.array_equal(got[keep].view(np.int32), want[r0:r1][keep].view(np.int32))
    none = kernels.glt_row_range(dev(np.zeros((4, 5), np.int32)), dev(np.zeros((4, 5), np.int32)), Hr, Wr).tolist()
    assert none[1] == 0


def test_row_window_fused_srf_and_u16_export():
    Hr, Wr, B = 120, 64, 285
    w = synthetic.emit_wavelengths()
    good = synthetic.good_band_mask(w)
    raw = synthetic.raw_cube_spectra_np((Hr, Wr, B), seed=2, good=good)
    gx, gy = synthetic.rotation_glt(Hr, Wr, 25.0)
    gx, gy = synthetic.inject_glt_defects(gx, gy, Hr, Wr, seed=4, n_oob=8, n_neg=8)
    W, names, _, fill_out = srf.srf_fold_weights(w, srf.synthetic_s2_srf(), good)
    Wd, fo = dev(W), dev(fill_out)
    Ho = gx.shape[0]
    fmf = torch.empty(gx.shape, dtype=torch.bool, device=DEV)
    bands, valid, diag, _ = kernels.glt_srf(dev(raw), dev(gx), dev(gy), Wd, fo, fit_mask_out=fmf, gate_k=0)
    q, qv, qb, qd = kernels.glt_ortho_u16(dev(raw), dev(gx), dev(gy))
    tot = torch.zeros(4, dtype=torch.int64, device=DEV)
    for rank in range(3):
        r0, r1 = hdist.shard_rows(Ho, rank, 3, align=8)
        sx, sy = dev(gx[r0:r1]), dev(gy[r0:r1])
        lo, hi = kernels.glt_row_range(sx, sy, Hr, Wr).tolist()
        assert hi - lo < Hr
        fm = torch.empty(sx.shape, dtype=torch.bool, device=DEV)
        b, v, d, o = kernels.glt_srf(dev(raw[lo:hi]), sx, sy, Wd, fo, fit_mask_out=fm, gate_k=0, raw_row0=lo,
                                     raw_rows_total=Hr, materialize_ortho=True)
        assert torch.equal(b.view(torch.int32), bands[:, r0:r1].contiguous().view(torch.int32))
        assert torch.equal(v, valid[r0:r1]) and torch.equal(fm, fmf[r0:r1])
        want, _, _ = oglt.glt_ortho(raw, gx[r0:r1], gy[r0:r1])
        assert np.array_equal(o.cpu().numpy().view(np.int32), want.view(np.int32))
        tot += d
        q2, qv2, qb2, _ = kernels.glt_ortho_u16(dev(raw[lo:hi]), sx, sy, raw_row0=lo, raw_rows_total=Hr)
        assert torch.equal(q2, q[:, r0:r1]) and torch.equal(qb2, qb[r0:r1]) and torch.equal(qv2, qv[r0:r1])
    assert tot.tolist() == diag.tolist() + [0]


@pytest.mark.parametrize("B", [285, 2])
def test_tile_batch_offsets_live_in_the_kernel(B):
    """A stack of independent tiles (tiles_helpers batch): tile t's GLT indexes ITS raw tile; an entry past its own
    tile is out of bounds (dropped and counted), never a read from the neighbour."""
    T, h, w = 5, 24, 40
    rng = np.random.default_rng(9)
    raw = synthetic.raw_cube_bits_np((T, h, w, B), seed=1)
    gx = rng.integers(0, w + 1, size=(T, h, w)).astype(np.int32)           # 0 = nodata, 1..w
    gy = rng.integers(0, h + 1, size=(T, h, w)).astype(np.int32)
    gy[rng.random((T, h, w)) < 0.02] = h + 1                               # just past the own tile: the next tile's row 0
    gy[rng.random((T, h, w)) < 0.01] = -2
    gx[rng.random((T, h, w)) < 0.01] = w + 3
    out, valid, diag = kernels.glt_ortho(dev(raw.reshape(T * h, w, B)), dev(gx.reshape(T * h, w)),
                                         dev(gy.reshape(T * h, w)), tile_rows=(h, h))
    nz = ib = 0
    for t in range(T):
        want, vref, dref = oglt.glt_ortho(raw[t], gx[t], gy[t])
        assert np.array_equal(out[t * h:(t + 1) * h].cpu().numpy().view(np.int32), want.view(np.int32)), f"tile {t}"
        assert np.array_equal(valid[t * h:(t + 1) * h].cpu().numpy(), vref)
        nz += dref["valid_glt_count"]
        ib += dref["valid_glt_inbounds_count"]
    assert diag.tolist() == [nz, ib, nz - ib, 0]
    if B >= 32:
        q, qv, _, _ = kernels.glt_ortho_u16(dev(raw.reshape(T * h, w, B)), dev(gx.reshape(T * h, w)),
                                            dev(gy.reshape(T * h, w)), tile_rows=(h, h))
        ref = kernels.quantize_u16(out.permute(2, 0, 1).contiguous(), nodata=-9999.0)
        assert torch.equal(q, ref) and torch.equal(qv, valid)


def test_synthesize_tiles_uses_the_glt_as_delivered():
    T, h, B = 6, 32, 285
    w = synthetic.emit_wavelengths()
    good = synthetic.good_band_mask(w)
    ps = PairSynthesizer(w, srf.synthetic_s2_srf(), good, deg=2, device=DEV, min_count=50)
    raw = np.stack([synthetic.raw_cube_spectra_np((h, h, B), seed=20 + t, good=good) for t in range(T)])
    gl = [synthetic.identity_glt(h, h, seed=2 + t) for t in range(T)]
    gx, gy = np.stack([g[0] for g in gl]), np.stack([g[1] for g in gl])
    gy[1, 3, 4] = h + 1                                                     # past the own tile
    rawd, gxd, gyd = dev(raw), dev(gx), dev(gy)
    b0 = torch.stack([ps.bands_from_raw(rawd[t], gxd[t], gyd[t])[0] for t in range(T)], dim=1)   # [K, T, h, h]
    s2 = synthetic.s2_reference_torch(b0, seed=3)
    res = ps.synthesize_tiles(rawd, gxd, gyd, s2)
    assert torch.equal(res.bands.view(torch.int32), b0.contiguous().view(torch.int32))
    assert not bool(res.valid[1, 3, 4])
    for t in range(T):
        one = ps.synthesize(rawd[t], gxd[t], gyd[t], s2[:, t].contiguous())
        assert torch.equal(one.fit_mask, res.fit_mask[t]) and torch.equal(one.valid, res.valid[t])
        torch.testing.assert_close(one.coeffs, res.coeffs[:, t], rtol=1e-9, atol=1e-12)
        assert (one.matched - res.matched[:, t]).abs().max().item() <= 1.2e-7


def test_moments_sum_is_the_ordered_sum():
    g = torch.Generator(device=DEV).manual_seed(1)
    per = [torch.rand((12, 8), generator=g, device=DEV, dtype=torch.float64) * 10.0 ** (3 * i) for i in range(7)]
    want = torch.zeros_like(per[0])
    for m in per:
        want += m
    assert torch.equal(kernels.moments_sum(per), want)
    assert torch.equal(kernels.moments_sum(torch.stack(per)), want)
    assert torch.equal(kernels.moments_sum([], like=per[0]), torch.zeros_like(per[0]))
    with pytest.raises(ValueError):
        kernels.moments_sum([])


def test_synthesize_slab_stages_only_the_rows_a_slab_references():
    """configs[4] on one GPU: the raw mosaic stays in HOST memory; every slab uploads its own band of raw rows."""
    Hr, Wr = 200, 180
    w = synthetic.emit_wavelengths()
    good = synthetic.good_band_mask(w)
    ps = PairSynthesizer(w, srf.synthetic_s2_srf(), good, deg=2, device=DEV)
    raw = synthetic.raw_cube_spectra_np((Hr, Wr, 285), seed=8, good=good)
    gx, gy = synthetic.rotation_glt(Hr, Wr, 25.0)
    gx, gy = synthetic.inject_glt_defects(gx, gy, Hr, Wr, seed=5, n_oob=8, n_neg=8)
    Ho = gx.shape[0]
    gxd, gyd, rawd = dev(gx), dev(gy), dev(raw)
    b0 = ps.bands_from_raw(rawd, gxd, gyd)[0]
    s2 = synthetic.s2_reference_torch(b0, seed=4)
    full = ps.synthesize(rawd, gxd, gyd, s2)
    host = torch.from_numpy(raw).pin_memory()
    stage = torch.empty(Hr * Wr * 285, dtype=torch.float32, device=DEV)
    staged_rows = 0
    tot = torch.zeros(3, dtype=torch.int64, device=DEV)
    for rank in range(4):
        r0, r1 = hdist.shard_rows(Ho, rank, 4, align=8)
        for src in (host, raw, rawd):                                       # pinned host tensor, numpy array, CUDA tensor
            res = ps.synthesize_slab(src, gxd[r0:r1], gyd[r0:r1], s2[:, r0:r1], stage=stage)
            assert torch.equal(res.bands.view(torch.int32), full.bands[:, r0:r1].contiguous().view(torch.int32))
            assert torch.equal(res.valid, full.valid[r0:r1]) and torch.equal(res.fit_mask, full.fit_mask[r0:r1])
            assert res.diag.numel() == 3
        lo, hi = res.raw_rows
        assert 0 <= lo < hi <= Hr and hi - lo < Hr      # 64 cos25 + Wo sin25 ~ 0.8 Hr rows: a band, not the cube
        staged_rows += hi - lo
        tot += res.diag
    assert torch.equal(tot, full.diag)
    assert staged_rows < 4 * Hr                                              # bands overlap, but nobody staged everything
    with pytest.raises(ValueError, match="stage holds"):
        ps.synthesize_slab(host, gxd[:8], gyd[:8], s2[:, :8], stage=stage[:10])

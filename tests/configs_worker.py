"""Worker of tests/test_multi_gpu_configs.py (torchrun, one process per GPU): BASELINE.json configs[3] and configs[4]
as written, at sizes the numpy oracle finishes in seconds.

configs[3]  64 synthetic granules, seeds 100..163, dealt round-robin (dist.shard_units); ONE global fit — through
            dist.PeerExchange (moments over NVLink peer memory) and through the NCCL all-reduce — then apply.
            Checked on every rank: coefficients against oracle.poly.polyfit_paired over the CONCATENATION of all 64
            granules (computed in this process with numpy), equal coefficients on every rank bit for bit, the rank's
            own matched planes against the oracle's apply.
configs[4]  one large ortho grid cut into row slabs (dist.shard_rows), raw mosaic kept in HOST memory, every rank
            stages only the raw rows its slab references (PairSynthesizer.synthesize_slab).  Checked: the slab's
            planes / masks / diagnostics equal the un-sharded single-GPU run's rows BIT FOR BIT, the staged window is
            a strict sub-range of the raw rows, the global fit equals the un-sharded fit.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hsr_b200 import dist as hdist  # noqa: E402
from hsr_b200 import synthetic  # noqa: E402
from hsr_b200.pipeline import PairSynthesizer  # noqa: E402
from hsr_b200.s2_emit.srf import synthetic_s2_srf  # noqa: E402
from oracle import glt as oglt  # noqa: E402
from oracle import poly as opoly  # noqa: E402
from oracle import srf as osrf  # noqa: E402

N_GRANULES = 64
SEED0 = 100
DEG = 2


def oracle_granule(seed, Hr, Wr, gx, gy, w, good, table, names):
    raw = synthetic.raw_cube_spectra_np((Hr, Wr, 285), seed=seed, good=good)
    ortho, valid, _ = oglt.glt_ortho(raw, gx, gy)
    ref = osrf.pseudo_s2_srf_integral(ortho, w, table, good)
    x = np.stack([ref[b] for b in names]).astype(np.float32)
    s2 = synthetic.s2_reference_np(x, seed=seed + 1000)
    fm = opoly.fit_mask(x, valid, 0, 0.0)
    return raw, x, s2, fm


def config3(rank, world, device, px):
    w = synthetic.emit_wavelengths()
    good = synthetic.good_band_mask(w)
    table = synthetic_s2_srf()
    ps = PairSynthesizer(w, table, good, deg=DEG, device=device)
    Hr, Wr = 40, 34
    gx, gy = synthetic.rotation_glt(Hr, Wr, 25.0)
    gxd, gyd = torch.from_numpy(gx).to(device), torch.from_numpy(gy).to(device)
    # the oracle over ALL granules (every rank computes it: the check is local, nothing is trusted from a peer)
    everything = [oracle_granule(SEED0 + i, Hr, Wr, gx, gy, w, good, table, ps.band_names) for i in range(N_GRANULES)]
    xcat = np.concatenate([e[1].reshape(ps.K, -1) for e in everything], axis=1)
    ycat = np.concatenate([e[2].reshape(ps.K, -1) for e in everything], axis=1)
    mcat = np.concatenate([e[3].reshape(-1) for e in everything])
    want = opoly.polyfit_paired(xcat, ycat, mcat, DEG, min_count=200)

    mine = hdist.shard_units(N_GRANULES, rank, world)
    assert mine == list(range(rank, N_GRANULES, world))
    granules = [{"raw": torch.from_numpy(everything[i][0]).to(device), "glt_x": gxd, "glt_y": gyd,
                 "s2_ref": torch.from_numpy(everything[i][2]).to(device)} for i in mine]
    for path in ("peer", "nccl"):
        res = ps.synthesize_sharded(granules, exchange=px if path == "peer" else None)
        torch.cuda.synchronize()
        if path == "peer":
            px.check()
        assert len(res) == len(mine)
        c = res[0].coeffs
        err = float(np.max(np.abs(c.cpu().numpy() - want).max(1) / np.abs(want).max(1)))
        assert err < 1e-4, f"{path}: global coefficients off by {err:.2e}"
        gathered = [torch.empty_like(c) for _ in range(world)]
        dist.all_gather(gathered, c.contiguous())
        if path == "peer":
            assert all(torch.equal(g, gathered[0]) for g in gathered), "ranks solved different systems"
        else:
            assert all(torch.allclose(g, gathered[0], rtol=1e-9, atol=1e-12) for g in gathered)
        for r, i in zip(res, mine):
            assert torch.equal(r.coeffs, c)
            _, x, _, fm = everything[i]
            assert np.array_equal(r.fit_mask.cpu().numpy(), fm)
            matched = opoly.apply_poly_planes(x, want, fm)
            assert np.max(np.abs(r.matched.cpu().numpy() - matched)) < 1e-4
        if path == "peer":
            peer_coeffs = c.clone()
        else:
            torch.testing.assert_close(c, peer_coeffs, rtol=1e-9, atol=1e-12)
    # uneven deal: 5 granules over the ranks (some ranks may own none) still gives one global fit and no hang
    few = [g for g, i in zip(granules, mine) if i < 5]
    res = ps.synthesize_sharded(few, exchange=px)
    torch.cuda.synchronize()
    px.check()
    want5 = opoly.polyfit_paired(np.concatenate([everything[i][1].reshape(ps.K, -1) for i in range(5)], axis=1),
                                 np.concatenate([everything[i][2].reshape(ps.K, -1) for i in range(5)], axis=1),
                                 np.concatenate([everything[i][3].reshape(-1) for i in range(5)]), DEG, min_count=200)
    for r in res:
        err = float(np.max(np.abs(r.coeffs.cpu().numpy() - want5).max(1) / np.abs(want5).max(1)))
        assert err < 1e-4, f"uneven deal: coefficients off by {err:.2e}"
    return True


def config4(rank, world, device, px):
    w = synthetic.emit_wavelengths()
    good = synthetic.good_band_mask(w)
    table = synthetic_s2_srf()
    ps = PairSynthesizer(w, table, good, deg=DEG, device=device)
    Hr, Wr = 300, 300                                   # raw mosaic (host memory), 25-degree GLT -> ~399 x 399 ortho grid
    raw = synthetic.raw_cube_spectra_np((Hr, Wr, 285), seed=7, good=good)
    gx, gy = synthetic.rotation_glt(Hr, Wr, 25.0)
    gx, gy = synthetic.inject_glt_defects(gx, gy, Hr, Wr, seed=3, n_oob=16, n_neg=16)
    Ho, Wo = gx.shape
    gxd, gyd = torch.from_numpy(gx).to(device), torch.from_numpy(gy).to(device)
    # the un-sharded run (every rank has the memory for it at this size) is the reference for "bit-equal"
    rawd = torch.from_numpy(raw).to(device)
    b0 = ps.bands_from_raw(rawd, gxd, gyd)[0]
    s2 = synthetic.s2_reference_torch(b0, seed=11)
    full = ps.synthesize(rawd, gxd, gyd, s2)
    torch.cuda.synchronize()
    del rawd
    raw_host = torch.from_numpy(raw).pin_memory()
    r0, r1 = hdist.shard_rows(Ho, rank, world, align=8)
    assert r1 > r0
    res = ps.synthesize_slab(raw_host, gxd[r0:r1], gyd[r0:r1], s2[:, r0:r1], exchange=px)
    torch.cuda.synchronize()
    px.check()
    lo, hi = res.raw_rows
    if world > 1:
        assert hi - lo < Hr, f"slab staged all {Hr} raw rows"
    assert torch.equal(res.bands.view(torch.int32), full.bands[:, r0:r1].contiguous().view(torch.int32)), "planes differ"
    assert torch.equal(res.valid, full.valid[r0:r1]) and torch.equal(res.fit_mask, full.fit_mask[r0:r1])
    tot = res.diag.clone()
    dist.all_reduce(tot)
    assert torch.equal(tot, full.diag), "GLT diagnostics of the slabs do not add up"
    # global fit over the slabs == the un-sharded fit (sums in a different order: equal to rounding), same planes
    torch.testing.assert_close(res.moments, full.moments, rtol=1e-12, atol=0)
    torch.testing.assert_close(res.coeffs, full.coeffs, rtol=1e-8, atol=1e-11)
    assert (res.matched - full.matched[:, r0:r1]).abs().max().item() <= 2.4e-7
    # a window that is too small must be reported, not silently filled
    bad = ps.bands_from_raw(torch.from_numpy(raw[lo + 1:hi]).to(device), gxd[r0:r1], gyd[r0:r1], raw_row0=lo + 1,
                            raw_rows_total=Hr)
    assert int(bad[2][3]) > 0, "entries outside the staged window were not counted"
    return True


def main():
    rank, world, device = hdist.init_from_env()
    px = hdist.PeerExchange(device=device, timeout_ms=20000)
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "config3"):
        config3(rank, world, device, px)
    if which in ("all", "config4"):
        config4(rank, world, device, px)
    dist.barrier()
    torch.cuda.synchronize()
    px.close()
    if rank == 0:
        print("multi-gpu configs OK", world, which)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Worker of tests/test_peer_exchange.py (launched by torchrun, one process per GPU): the global fit through
dist.PeerExchange (moments over NVLink peer memory, fused into the finalize / solve kernels) must equal the NCCL
all-reduce path on every rank, for many consecutive epochs and with ranks running out of step."""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hsr_b200 import dist as hdist  # noqa: E402
from hsr_b200 import kernels, synthetic  # noqa: E402
from hsr_b200.pipeline import PairSynthesizer  # noqa: E402
from hsr_b200.s2_emit.srf import synthetic_s2_srf  # noqa: E402


def main():
    rank, world, device = hdist.init_from_env()
    w = synthetic.emit_wavelengths()
    good = synthetic.good_band_mask(w)
    ps = PairSynthesizer(w, synthetic_s2_srf(), good, deg=2, device=device)
    px = hdist.PeerExchange(device=device)
    assert px.world == world and px.rank == rank
    Hr, Wr = 120, 90
    gx_np, gy_np = synthetic.rotation_glt(Hr, Wr, 25.0)
    gx, gy = torch.from_numpy(gx_np).to(device), torch.from_numpy(gy_np).to(device)
    for epoch in range(7):
        raw = synthetic.raw_cube_spectra_torch((Hr, Wr, 285), seed=1000 * epoch + rank, device=device, good=good)
        bands0 = ps.bands_from_raw(raw, gx, gy)[0]
        s2 = synthetic.s2_reference_torch(bands0, seed=7 * epoch + rank)
        if (epoch + rank) % 2 == 0:
            time.sleep(0.05)                       # ranks out of step: the flags, not luck, must order the exchange
        a = ps.synthesize(raw, gx, gy, s2, exchange=px)
        b = ps.synthesize(raw, gx, gy, s2, allreduce=True)
        torch.cuda.synchronize()
        # same global sums (rank-ordered vs NCCL's order: equal to rounding), same fit, same planes
        torch.testing.assert_close(a.moments, b.moments, rtol=1e-14, atol=0)
        torch.testing.assert_close(a.coeffs, b.coeffs, rtol=1e-10, atol=1e-13)
        assert (a.matched - b.matched).abs().max().item() <= 1.2e-7
        # every rank solved the SAME system bit for bit
        gathered = [torch.empty_like(a.moments) for _ in range(world)]
        dist.all_gather(gathered, a.moments.contiguous())
        assert all(torch.equal(g, gathered[0]) for g in gathered)
        cg = [torch.empty_like(a.coeffs) for _ in range(world)]
        dist.all_gather(cg, a.coeffs.contiguous())
        assert all(torch.equal(c, cg[0]) for c in cg)
        # and it is the sum of the local moments
        local = kernels.fit_moments(a.bands, s2, a.fit_mask, 2, mask_given=True)[0].view(ps.K, -1)
        tot = local.clone()
        dist.all_reduce(tot)
        torch.testing.assert_close(a.moments, tot, rtol=1e-14, atol=0)
        assert not torch.equal(a.moments, local) or world == 1
    dist.barrier()
    torch.cuda.synchronize()
    px.close()
    if rank == 0:
        print("peer exchange OK", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

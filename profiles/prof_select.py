"""Timing of the exact masked percentiles (3 launches, one pass) on granule-sized planes, K = 3 and 12, single set and
the (x, y) pair of the shared stretch; how many series fell back to the whole-series select.
    python profiles/prof_select.py [reps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hsr_b200 import kernels  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = "cuda"
n = 1685 * 1667
g = torch.Generator(device=dev).manual_seed(0)
mask = torch.rand(n, generator=g, device=dev) < 0.566


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


SEL_STATE_BYTES = 96        # sizeof(SelSeries): 2+2 u32 keys, 2 u32 counts, fell_back, pad..., see csrc/select.cu
for K in (3, 12):
    for name, gen in (("uniform random (adversarial)", lambda: torch.rand((K, n), generator=g, device=dev)),
                      ("image-like (smooth + noise, 10 % exact zeros)",
                       lambda: (torch.rand((K, n), generator=g, device=dev) ** 2 * 0.8).clamp_(0.08, 1.0) - 0.08)):
        x = kernels.alloc_planes(K, (n,), dev)
        x.copy_(gen())
        y = kernels.alloc_planes(K, (n,), dev)
        y.copy_(gen())
        ms1 = timed(lambda: kernels.masked_percentiles(x, mask, [2, 98]))
        ms2 = timed(lambda: kernels.masked_percentiles(x, mask, [2, 98], y=y))
        ws = []
        ox, oy = kernels.masked_percentiles(x, mask, [2, 98], y=y, _workspace_out=ws)
        torch.cuda.synchronize()
        st = ws[0][:2 * K * SEL_STATE_BYTES].cpu().numpy().view(np.uint32).reshape(2 * K, -1)
        ref = np.percentile(x[0].cpu().numpy()[mask.cpu().numpy()], [2, 98])
        ok = np.array_equal(ox[0, 0].cpu().numpy(), ref)
        gb1, gb2 = (K * n * 4 + n) / 1e9, (2 * K * n * 4 + n) / 1e9
        print(f"K={K:2d} {name:48s} single {ms1:7.3f} ms ({gb1 / ms1 * 1e3:6.0f} GB/s)   pair {ms2:7.3f} ms "
              f"({gb2 / ms2 * 1e3:6.0f} GB/s)   candidates/bracket ~{int(st[:, 4:6].mean())}   fell back: {int(st[:, 6].sum())}/{2 * K}"
              f"   exact vs numpy: {ok}")

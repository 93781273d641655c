"""tcgen05.mma issue-rate probe (csrc/probe.cu): cycles per M128 x N x K16 instruction, SS vs TS operands, and how the
rate depends on whether consecutive instructions accumulate into the same TMEM region (dmode 0) or rotate over 2 / 3."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nu_nerf_b200._lib import call  # noqa: E402

iters = 256
grid = 148
for mode, name in ((0, "SS"), (1, "TS")):
    for N in (64, 128, 256):
        for dmode in (0, 1, 2):
            if dmode and N > 128:
                continue
            out = torch.zeros(2 * grid, dtype=torch.int64, device="cuda")
            for _ in range(2):
                call("nunerf_mma_probe", mode, N, iters, grid, 4, dmode, out.data_ptr())
            torch.cuda.synchronize()
            o = out.view(grid, 2).float()
            print(f"{name} N={N:3d} dmode {dmode}: issue {o[:, 0].mean().item() / (4 * iters):7.1f} cyc/MMA, "
                  f"complete {o[:, 1].mean().item() / (4 * iters):7.1f} cyc/MMA (max CTA {o[:, 1].max().item() / (4 * iters):7.1f}); "
                  f"floor {N / 2:.0f}", flush=True)

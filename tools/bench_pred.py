"""Micro-benchmark of one make_predictor MLP (field.py:371-408), forward and dX-chain backward, fused (csrc/chain.cu)
against layer by layer (csrc/linear.cu): CUDA events, algorithmic FLOPs = 2 * M * sum(N * K) of the layers run."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nu_nerf_b200 import engine as eng  # noqa: E402
from nu_nerf_b200.renderer_zerothick import NeROShapeRenderer, load_default_cfg  # noqa: E402
from bench_chain import timeit  # noqa: E402


def main():
    cfg = load_default_cfg()
    cfg["precision"] = "bf16"
    torch.manual_seed(0)
    net = NeROShapeRenderer(cfg, training=False).cuda()
    w = net._prepare()
    pw = w.pred["outer_light"]
    if len(sys.argv) > 1 and sys.argv[1] == "once":          # two launches only (ncu --set full captures)
        M = 383_000
        x = eng.P(M, 128, 1, "cuda", zero=True)
        x.t[:, :72] = torch.randn(M, 72, device="cuda").to(torch.bfloat16)
        dz = eng.P(M, 64, 1, "cuda", zero=True)
        dz.t[:, :3] = torch.randn(M, 3, device="cuda").to(torch.bfloat16)
        dx = torch.empty(M, 128, device="cuda")
        for _ in range(2):
            t = eng.pred_forward(pw, x, M, 128, 1)
            eng.pred_backward(pw, t, dz, 1, dx_f32=dx, dx_n=128)
        torch.cuda.synchronize()
        print("ok")
        return
    for M in (383_000, 3 * 383_000):
        x = eng.P(M, 128, 1, "cuda", zero=True)
        x.t[:, :72] = torch.randn(M, 72, device="cuda").to(torch.bfloat16)
        dz = eng.P(M, 64, 1, "cuda", zero=True)
        dz.t[:, :3] = torch.randn(M, 3, device="cuda").to(torch.bfloat16)
        dx = torch.empty(M, 128, device="cuda")
        f_fwd = 2.0 * M * (128 * 256 + 2 * 256 * 256 + 256 * 16)
        f_bwd = 2.0 * M * (64 * 256 + 2 * 256 * 256 + 256 * 128)
        f_dw = 2.0 * M * (3 * 256 + 2 * 256 * 256 + 256 * 128)
        for fused in (True, False):
            eng.FUSED_CHAINS = fused
            t = eng.pred_forward(pw, x, M, 128, 1)
            ms_f = timeit(lambda: eng.pred_forward(pw, x, M, 128, 1))
            ms_b = timeit(lambda: eng.pred_backward(pw, t, dz, 1, dx_f32=dx, dx_n=128))
            name = "fused" if fused else "layerwise"
            print(f"predictor M={M:8d} {name:10s} fwd {ms_f*1e3:8.1f} us {f_fwd/ms_f/1e9:7.1f} TFLOP/s | "
                  f"bwd (dX chain + 4 dW) {ms_b*1e3:8.1f} us {(f_bwd+f_dw)/ms_b/1e9:7.1f} TFLOP/s", flush=True)
        eng.FUSED_CHAINS = True


if __name__ == "__main__":
    main()

"""Micro-benchmark of the fused SDF-inference chain (csrc/chain.cu) against the layer-by-layer path (CUDA events)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nu_nerf_b200 import engine as eng  # noqa: E402
from nu_nerf_b200.renderer_zerothick import NeROShapeRenderer, load_default_cfg  # noqa: E402

M_SDF_HEAD = 459008


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    cfg = load_default_cfg()
    cfg["precision"] = "bf16"
    torch.manual_seed(0)
    net = NeROShapeRenderer(cfg, training=False).cuda()
    w = net._prepare()
    for M in (4096 * 64, 4096 * 112, 4 * 1024 * 1024):
        pts = (torch.rand(M, 3, device="cuda") * 2 - 1).contiguous()
        for name, fused in (("fused chain", True), ("layer by layer", False)):
            ms = timeit(lambda: eng.sdf_infer(w.sdf, pts, 1, fused=fused))
            print(f"sdf_infer M={M:8d} {name:15s} {ms*1e3:9.1f} us  {2.0*M_SDF_HEAD*M/ms/1e9:7.1f} TFLOP/s  "
                  f"{M/ms/1e3:7.1f} Mpts/s", flush=True)


if __name__ == "__main__":
    main()

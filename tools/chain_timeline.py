"""Timeline of one tile of the fused chain kernel (clock64 stamps of CTA 0): when each 64-column chunk of the activation
was published by the epilogue warps, when the MMA warp saw it and when it finished issuing the k-block."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nu_nerf_b200 import engine as eng  # noqa: E402
from nu_nerf_b200.renderer_zerothick import NeROShapeRenderer, load_default_cfg  # noqa: E402

cfg = load_default_cfg()
cfg["precision"] = "bf16"
torch.manual_seed(0)
net = NeROShapeRenderer(cfg, training=False).cuda()
w = net._prepare()
M = 148 * 128 * 4
pts = (torch.rand(M, 3, device="cuda") * 2 - 1).contiguous()
tl = torch.zeros(512, dtype=torch.int64, device="cuda")
for _ in range(3):
    eng.sdf_infer_fused(w.sdf, pts, timeline=tl)
torch.cuda.synchronize()
t = tl.cpu().tolist()
t0 = min(x for x in t if x > 0)
print("layer: epilogue start (warp0 / warp15) | chunk publish times (warp 0) | MMA: saw chunk c / issued k-block c")
for l in range(9):
    e0 = [t[256 + l * 8 + i] - t0 if t[256 + l * 8 + i] else -1 for i in range(5)]
    e15 = [t[256 + 128 + l * 8 + i] - t0 if t[256 + 128 + l * 8 + i] else -1 for i in range(5)]
    mm = [(t[(l * 4 + b) * 2] - t0 if t[(l * 4 + b) * 2] else -1, t[(l * 4 + b) * 2 + 1] - t0 if t[(l * 4 + b) * 2 + 1] else -1)
          for b in range(4)]
    print(f"L{l}: epi0 {e0}  epi15 {e15}  mma {mm}")

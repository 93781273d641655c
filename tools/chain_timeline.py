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
rel = lambda v: v - t0 if v else -1
print("layer/tile: MMA issued+committed at | epilogue start -> end   (cycles since the first stamp, second tile pair of CTA 0)")
for l in range(9):
    for tt in range(2):
        i = (l * 2 + tt) * 2
        print(f"L{l} tile{tt}: mma commit {rel(t[i]):7d} | epilogue {rel(t[256 + i]):7d} -> {rel(t[256 + i + 1]):7d}")

print("layer 5 detail: producer saw w_empty (tile, kb) | MMA: x_done seen, per kb: w_full seen -> issued")
for tt in range(2):
    print(f" tile{tt}: x_done wait begins {rel(t[450 + tt])} seen {rel(t[420 + tt])}")
    for kb in range(4):
        print(f"   kb{kb}: producer w_empty seen {rel(t[400 + tt * 4 + kb]):7d} | w_full seen {rel(t[430 + (tt * 4 + kb) * 2]):7d} -> issued {rel(t[430 + (tt * 4 + kb) * 2 + 1]):7d}")

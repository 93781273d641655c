"""Timeline of the second tile of CTA 0 in the TS-mode chain kernel (csrc/chain_ts.cu, clock64 stamps): per layer and
K-block when the MMA warp had its operands and when it had issued; per accumulator half when it was committed, when the
epilogue saw it, had read it, and had published each chunk of the next A operand."""
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
for l in range(9):
    kbs = " ".join(f"kb{kb}:{rel(t[64 + (l * 4 + kb) * 2])}->{rel(t[64 + (l * 4 + kb) * 2 + 1])}" for kb in range(4))
    print(f"L{l} MMA operands ready -> issued: {kbs}")
    for h in range(2):
        i = (l * 2 + h) * 2
        print(f"   half{h}: commit issued {rel(t[i]):7d} | epilogue saw d_full {rel(t[256 + i]):7d}, read {rel(t[320 + l * 4 + 2 * h]):7d}/{rel(t[320 + l * 4 + 2 * h + 1]):7d}, "
              f"chunks published {rel(t[360 + l * 4 + 2 * h]):7d} {rel(t[360 + l * 4 + 2 * h + 1]):7d}, done {rel(t[256 + i + 1]):7d}"
              f" | math done / st complete: {rel(t[400 + (l * 4 + 2 * h) * 2])}/{rel(t[400 + (l * 4 + 2 * h) * 2 + 1])} "
              f"{rel(t[400 + (l * 4 + 2 * h + 1) * 2])}/{rel(t[400 + (l * 4 + 2 * h + 1) * 2 + 1])}")

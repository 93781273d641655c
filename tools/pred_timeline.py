"""Per-layer clock64 timeline of the fused predictor forward chain (CTA 0, second tile pair)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nu_nerf_b200 import engine as eng  # noqa: E402
from nu_nerf_b200.ops import chain  # noqa: E402
from nu_nerf_b200.renderer_zerothick import NeROShapeRenderer, load_default_cfg  # noqa: E402

cfg = load_default_cfg()
cfg["precision"] = "bf16"
torch.manual_seed(0)
net = NeROShapeRenderer(cfg, training=False).cuda()
w = net._prepare()
pw = w.pred["outer_light"]
M = 383_000
x = eng.P(M, 128, 1, "cuda", zero=True)
x.t[:, :72] = torch.randn(M, 72, device="cuda").to(torch.bfloat16)
H = [eng.P(M, 256, 1, "cuda") for _ in range(3)]
Mk = [torch.empty(M, 32, dtype=torch.uint8, device="cuda") for _ in range(3)]
head = torch.empty(M, 16, device="cuda")
hidden = lambda i, K: dict(W=pw.L[i].Wk, N=256, K=K, bias=pw.L[i].b, act=1, mask_out=Mk[i], store=H[i], keep=1)
lays = [hidden(0, 128), hidden(1, 256), hidden(2, 256), dict(W=pw.L[3].Wk, N=16, K=256, bias=pw.L[3].b, out32=head, n32=16)]
tl = torch.zeros(512, dtype=torch.int64, device="cuda")
for _ in range(3):
    tl.zero_()
    chain(x, M, 128, lays, timeline=tl)
torch.cuda.synchronize()
t = tl.cpu().tolist()
t0 = min(v for v in t if v > 0)
rel = lambda v: v - t0 if v else -1
print("layer/tile: MMA committed at | epilogue start -> end (cycles, second tile pair of CTA 0)")
for l in range(4):
    for tt in range(2):
        i = (l * 2 + tt) * 2
        print(f"L{l} tile{tt}: mma commit {rel(t[i]):7d} | epilogue {rel(t[256 + i]):7d} -> {rel(t[256 + i + 1]):7d}")

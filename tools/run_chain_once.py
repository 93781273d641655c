"""One fused sdf_infer launch (for ncu captures)."""
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
M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096 * 112
pts = (torch.rand(M, 3, device="cuda") * 2 - 1).contiguous()
for _ in range(3):
    out = eng.sdf_infer(w.sdf, pts, 1, fused=True)
torch.cuda.synchronize()
print("ok", float(out.sum()))

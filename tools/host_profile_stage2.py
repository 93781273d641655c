"""Synchronising calls of one stage-2 training step (config-4 mesh, 4096 rays), with their Python locations."""
import collections
import os
import sys
import time
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import make_stage2, uv_sphere  # noqa: E402

R = 4096
V, Fc = uv_sphere(0.6, 224, 224)
net = make_stage2("bf16", mesh=(V, Fc)).cuda()
g = torch.Generator().manual_seed(1)
o = 3.0 * torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1)
d = torch.nn.functional.normalize(-o + 0.3 * torch.randn(R, 3, generator=g), dim=-1)
o, d = o.cuda(), d.cuda()
gt = torch.rand(R, 3, generator=g).cuda()
opt = torch.optim.Adam([p for p in net.parameters() if p.requires_grad], lr=5e-4, fused=True)


def step():
    opt.zero_grad(set_to_none=True)
    out = net.render(o, d, None, None, None, -1, 0.2, is_train=True, step=10000, is_nerf=True)
    tm = out["tir_mask"].detach()
    loss = net.compute_rgb_loss(out["ray_rgb"] * tm, gt * tm).mean() + (0.02 * out["gradient_error"]).mean()
    loss.backward()
    opt.step()
    return loss


for _ in range(8):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    step()
torch.cuda.synchronize()
print(f"stage-2 train step: {(time.perf_counter() - t0) * 100:.2f} ms")
torch.cuda.set_sync_debug_mode("warn")
with warnings.catch_warnings(record=True) as ws:
    warnings.simplefilter("always")
    step()
torch.cuda.set_sync_debug_mode("default")
cnt = collections.Counter(f"{w.filename.split('/repo/')[-1]}:{w.lineno}" for w in ws)
print(len(ws), "synchronising calls in one step")
for k, v in cnt.most_common(40):
    print(f"{v:4d}x {k}")

"""torch.profiler kernel table of one stage-2 training step (config 4 mesh, 4096 rays).  `--thick`: the non-zero-thickness
renderer of network/renderer.py (nu_nerf_b200/renderer.py)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import make_stage2, uv_sphere
R = 4096
V, Fc = uv_sphere(0.6, 224, 224)
THICK = "--thick" in sys.argv
net = make_stage2("bf16", mesh=(V, Fc), thick=THICK).cuda()
g = torch.Generator().manual_seed(1)
o = 3.0 * torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1)
d = torch.nn.functional.normalize(-o + 0.3 * torch.randn(R, 3, generator=g), dim=-1)
o, d = o.cuda(), d.cuda()
gt = torch.rand(R, 3, generator=g).cuda()
def step():
    net.zero_grad(set_to_none=True)
    args = (o, d, None, None, None, None) if THICK else (o, d, None, None, None)      # NZ:1482 carries a mask argument
    out = net.render(*args, -1, 0.2, is_train=True, step=10000, is_nerf=True)
    tm = out["tir_mask"].detach()
    loss = net.compute_rgb_loss(out["ray_rgb"] * tm, gt * tm).mean() + (0.02 * out["gradient_error"]).mean()
    loss.backward()
for _ in range(3): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(5):
    step()
torch.cuda.synchronize()
wall_ms = (time.perf_counter() - t0) / 5 * 1e3
ev = prof.key_averages()
kern = sum(e.self_device_time_total for e in ev)
syncs = {k: sum(e.count for e in ev if e.key == k) for k in ("aten::nonzero", "aten::item", "aten::_local_scalar_dense",
                                                             "cudaStreamSynchronize", "cudaMemcpyAsync", "cudaLaunchKernel")}
print(f"wall per step {wall_ms:.2f} ms; sum of kernel (self device) time {kern / 1e3:.2f} ms; host-side counts {syncs}")
rows = sorted(ev, key=lambda e: -e.device_time_total)[:28]
tot = sum(e.device_time_total for e in ev)
print("total device us", tot)
for e in rows:
    print(f"{e.device_time_total:10.0f} us {e.count:5d}x  {e.key[:90]}")

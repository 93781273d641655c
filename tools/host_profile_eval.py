"""Synchronising calls and wall time of one stage-1 eval render (is_train=False: depth, normal, shading buffers,
occ_prob_gt) of 4096 rays."""
import collections
import os
import sys
import time
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nu_nerf_b200.renderer_zerothick import NeROShapeRenderer, load_default_cfg
from nu_nerf_b200 import synthetic as syn
cfg = load_default_cfg(); cfg["precision"] = "bf16"
torch.manual_seed(0)
net = NeROShapeRenderer(cfg, training=False).cuda()
R = 4096
o, d = (t.cuda() for t in syn.synthetic_rays(R))
near, far = net.near_far_from_sphere(o, d)
def run():
    with torch.no_grad():
        return net.render(o, d, near, far, None, 0, 0.2, is_train=False, step=10000, is_nerf=False)
for _ in range(5): run()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): run()
torch.cuda.synchronize(); print(f"eval render {(time.perf_counter()-t0)*100:.2f} ms")
torch.cuda.set_sync_debug_mode("warn")
with warnings.catch_warnings(record=True) as ws:
    warnings.simplefilter("always"); run()
torch.cuda.set_sync_debug_mode("default")
cnt = collections.Counter(f"{w.filename.split('/repo/')[-1]}:{w.lineno}" for w in ws)
print(len(ws), "syncs"); [print(v, k) for k, v in cnt.most_common(20)]

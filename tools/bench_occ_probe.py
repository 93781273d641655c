"""Micro-benchmark of the occlusion-probe path (ZT:695-723): candidate selection primitives on 383k samples and
occl. probability of 2048 probe rays (points + fused SDF query + probe_weights_kernel, twice)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dev = "cuda"
M=383000
keys=torch.rand(M,device=dev)
def t(fn,n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n*1e3
print("topk 2048 of 383k unsorted: %.0f us"%t(lambda: torch.topk(keys,2048,largest=False,sorted=False)))
print("topk 2048 sorted: %.0f us"%t(lambda: torch.topk(keys,2048,largest=False)))
print("sort: %.0f us"%t(lambda: torch.sort(keys)))
print("rand: %.0f us"%t(lambda: torch.rand(M,device=dev)))
mask=keys<0.05
print("nonzero: %.0f us"%t(lambda: torch.nonzero(mask)))
print("cumsum: %.0f us"%t(lambda: torch.cumsum(mask.int(),0)))
from nu_nerf_b200.renderer_zerothick import NeROShapeRenderer, load_default_cfg
cfg=load_default_cfg(); cfg["precision"]="bf16"; torch.manual_seed(0)
net=NeROShapeRenderer(cfg,training=False).cuda(); w=net._prepare()
pts=torch.nn.functional.normalize(torch.randn(2048,3,device=dev),dim=-1)*0.5
dirs=torch.nn.functional.normalize(torch.randn(2048,3,device=dev),dim=-1)
print("occ_probability(2048): %.0f us"%t(lambda: net.occ_probability(pts,dirs,w)))

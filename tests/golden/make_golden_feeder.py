"""Generates tests/golden/feeder.npz with the UNMODIFIED reference's ray-table code (ZT:193-260, :347-361) on a tiny
synthetic image set.  Run in the build container only:  python tests/golden/make_golden_feeder.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    net, _ = rh.load_stage1(seed=0)
    g = torch.Generator().manual_seed(11)
    imn, h, w = 3, 4, 5
    imgs = torch.rand(imn, 3, h, w, generator=g)
    masks = (torch.rand(imn, h, w, generator=g) > 0.5).float()
    Ks = torch.tensor([[20.0, 0, 2.5], [0, 22.0, 2.0], [0, 0, 1]]).repeat(imn, 1, 1)
    rot = torch.linalg.qr(torch.randn(imn, 3, 3, generator=g))[0]
    poses = torch.cat([rot, torch.randn(imn, 3, 1, generator=g)], -1)
    info = {"imgs": imgs, "Ks": Ks, "poses": poses, "masks": masks}
    with rh.in_ref_dir():
        nb, _, rn, _, _ = net._construct_nerf_ray_batch(info)
        rb, _, rn2, _, _ = net._construct_ray_batch(info)
        # get_human_coordinate_poses (ZT:328-345) writes in place into an expanded tensor, which current torch rejects;
        # it does not touch the rays, so it is bypassed here (human_light is off in every shipped config)
        net.get_human_coordinate_poses = lambda p_: p_
        ro, rd, near, far, _ = net._process_ray_batch(rb, poses)
    res = {"imgs": imgs, "masks": masks, "Ks": Ks, "poses": poses}
    for k, v in nb.items():
        res["nerf_" + k] = v
    for k, v in rb.items():
        res["plain_" + k] = v
    res["plain_rays_o"], res["plain_rays_d"] = ro, rd
    # shuffle + two training slices exactly as ZT:193-197, :449-453 do (CPU generator seeded with 7)
    torch.manual_seed(7)
    perm = torch.randperm(rn, device="cpu")
    res["perm"] = perm
    res["slice0_rays_o"] = nb["rays_o"][perm][0:8]
    res["slice1_rgbs"] = nb["rgbs"][perm][8:16]
    np.savez_compressed(os.path.join(OUT, "feeder.npz"), **{k: v.numpy() for k, v in res.items()})
    print("feeder.npz", os.path.getsize(os.path.join(OUT, "feeder.npz")), "rays", rn)


if __name__ == "__main__":
    main()

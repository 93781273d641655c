"""Device-side ray-batch feeder: the ray-table part of NeROShapeRenderer._init_dataset / train_step
(network/renderer_zerothick.py, "ZT") kept resident in HBM.

  construct_nerf_ray_batch  ZT:222-254   per-pixel rays of every training view (is_nerf datasets: OpenGL camera axes)
  construct_ray_batch       ZT:199-220   pixel directions K^-1 [u+.5, v+.5, 1] + image index (world rays at fetch time,
                                         ZT:347-361)
  shuffle                   ZT:193-197   one permutation of the whole table
  __call__(step, n)         ZT:449-453   the next n rays: a device gather through the permutation instead of a host
                                         slice + H2D copy; reshuffles when fewer than 2n rays remain, like the reference

At the rates of this engine (> 1e5 rays/s/GPU) the reference's per-step `v[i:i+rn].cuda()` on a host-resident table
costs more than the render step's launch overhead; here a step costs one index slice and one gather per key.
Data-parallel ranks draw the SAME permutation (same generator seed) and take rank-strided elements of each global batch
(nu_nerf_b200.dist.shard_batch), which keeps the single-process semantics.
The image database itself (reading files, masks, intrinsics) stays outside the hot path: the feeder takes tensors.
"""
import torch
import torch.nn.functional as F


def construct_nerf_ray_batch(imgs, Ks, poses, masks=None):
    """imgs [imn,3,h,w], Ks [imn,3,3] (only Ks[0] is used, as in the reference), poses [imn,3,4] (camera-to-world).
    Returns ({'rgbs','idxs','rays_o','rays_d'[,'masks']}, rn, h, w) on the device of `imgs` (ZT:222-254)."""
    imn, _, h, w = imgs.shape
    dev = imgs.device
    i, j = torch.meshgrid(torch.linspace(0, w - 1, w, device=dev), torch.linspace(0, h - 1, h, device=dev), indexing="ij")
    i, j = i.t(), j.t()
    K = Ks[0].to(dev)
    dirs = torch.stack([(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], -torch.ones_like(i)], -1)     # h,w,3
    poses = poses.to(dev)
    # rays_d[n] = sum_k dirs[..., k] * R_n[:, k]   (ZT:238)
    rays_d = torch.sum(dirs[None, :, :, None, :] * poses[:, None, None, :3, :3], -1).reshape(imn, h * w, 3)
    rays_o = poses[:, None, :3, -1].expand(imn, h * w, 3)
    rn = imn * h * w
    batch = {
        "rgbs": imgs.permute(0, 2, 3, 1).reshape(rn, 3).float().contiguous(),
        "idxs": torch.arange(imn, dtype=torch.int64, device=dev)[:, None, None].repeat(1, h * w, 1).reshape(rn, 1),
        "rays_o": rays_o.reshape(rn, 3).float().contiguous(),
        "rays_d": rays_d.reshape(rn, 3).float().contiguous(),
    }
    if masks is not None:
        batch["masks"] = masks.reshape(rn).float().to(dev)
    return batch, rn, h, w


def construct_ray_batch(imgs, Ks):
    """Pixel-centre directions in camera space + image index (ZT:199-220)."""
    imn, _, h, w = imgs.shape
    dev = imgs.device
    ys, xs = torch.meshgrid(torch.arange(h, device=dev), torch.arange(w, device=dev), indexing="ij")
    coords = torch.stack([xs, ys], -1).float()[None].repeat(imn, 1, 1, 1).reshape(imn, h * w, 2)
    coords = torch.cat([coords + 0.5, torch.ones(imn, h * w, 1, device=dev)], 2)
    dirs = coords @ torch.inverse(Ks.to(dev)).permute(0, 2, 1)
    rn = imn * h * w
    batch = {
        "dirs": dirs.float().reshape(rn, 3).contiguous(),
        "rgbs": imgs.permute(0, 2, 3, 1).reshape(rn, 3).float().contiguous(),
        "idxs": torch.arange(imn, dtype=torch.int64, device=dev)[:, None, None].repeat(1, h * w, 1).reshape(rn, 1),
    }
    return batch, rn, h, w


def world_rays(dirs, idxs, poses):
    """_process_ray_batch (ZT:347-361) without the human-pose part: world-to-camera poses [imn,3,4] -> rays_o, rays_d."""
    idx = idxs[..., 0]
    rays_o = (poses[:, :, :3].permute(0, 2, 1) @ -poses[:, :, 3:])[idx, :, 0]
    rays_d = (poses[idx, :, :3].permute(0, 2, 1) @ dirs.unsqueeze(-1))[..., 0]
    return rays_o, F.normalize(rays_d, dim=-1)


class DeviceRayFeeder:
    """callable(step, n) -> {'rays_o','rays_d','rgbs'[, 'masks','idxs']} for `renderer.set_ray_source`."""

    def __init__(self, batch, poses=None, rank=0, world=1, seed=0, perm_device=None):
        self.batch = batch
        self.poses = poses                    # only for 'dirs' tables (non-nerf datasets): world-to-camera poses
        self.tbn = next(iter(batch.values())).shape[0]
        self.device = next(iter(batch.values())).device
        self.rank, self.world = rank, world
        # perm_device='cpu' draws the permutation exactly like the reference (torch.randperm on the CPU generator)
        self.perm_device = perm_device if perm_device is not None else self.device
        self.gen = torch.Generator(device=self.perm_device)
        self.gen.manual_seed(seed)
        self.shuffle()

    def shuffle(self):
        self.i = 0
        self.perm = torch.randperm(self.tbn, generator=self.gen, device=self.perm_device).to(self.device)

    def __call__(self, step, n):
        g = n * self.world                                     # global batch of this step
        idx = self.perm[self.i:self.i + g][self.rank::self.world]
        self.i += g
        if self.i + g >= self.tbn:                             # ZT:452
            self.shuffle()
        out = {k: v[idx] for k, v in self.batch.items()}
        if "dirs" in out:
            out["rays_o"], out["rays_d"] = world_rays(out.pop("dirs"), out["idxs"], self.poses)
        return out


def image_eval_source(imgs, Ks, poses, is_nerf=True, depths=None, masks=None):
    """Eval source for `renderer.set_eval_source` from image tensors (test_step, ZT:397-411): imgs [imn,3,h,w] at the
    evaluation resolution, Ks [imn,3,3], poses [imn,3,4] (camera-to-world for is_nerf datasets, world-to-camera
    otherwise), optional depths / masks [imn,h,w].  Returns fn(index) -> the rays and ground truth of that view."""
    def fn(index):
        sl = slice(index, index + 1)
        if is_nerf:
            batch, rn, h, w = construct_nerf_ray_batch(imgs[sl], Ks[sl], poses[sl])
            rays_o, rays_d = batch["rays_o"], batch["rays_d"]
        else:
            batch, rn, h, w = construct_ray_batch(imgs[sl], Ks[sl])
            rays_o, rays_d = world_rays(batch["dirs"], batch["idxs"], poses[sl].to(imgs.device))
        out = {"rays_o": rays_o, "rays_d": rays_d, "rgbs": batch["rgbs"], "h": h, "w": w}
        if depths is not None:
            out["gt_depth"] = depths[index]
        if masks is not None:
            out["gt_mask"] = masks[index].to(torch.int32)
        return out
    return fn

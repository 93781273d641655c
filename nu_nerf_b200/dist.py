"""Ray-sharded data parallelism for the stage-1 / stage-2 training step.

One process per GPU (torchrun); rays are independent units, the field (<= 27 MB of parameters) is replicated.
The reference trainer is single-GPU (train/trainer_zero.py): it shuffles the whole ray table once
(ZT:193-197) and slices `train_ray_num` rays per step (ZT:451-453).  Here every rank takes a rank-strided slice of the
same global batch, and the losses are normalised with GLOBAL denominators so that the SUM all-reduce of the flat
gradient equals the gradient of the single-process large-batch step:

    loss = sum_local(loss_rgb) / R_global + w_eik * sum_local(gradient_error) / N_in_global (+ ...)

`gradient_error` has a data-dependent length (the number of samples inside the unit sphere), so N_in_global needs one
scalar all-reduce before the backward pass; averaging per-rank means would weight ranks unequally.

Communication: one all-reduce of one flat fp32 buffer per step (NCCL over NVLink/NVSwitch on GPUs; the same code runs
on gloo for the CPU tests).  No other collective is on the path.
"""
import math

import torch
import torch.distributed as dist


def world_info(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_batch(indices, rank, world):
    """Rank-strided slice of a global batch index vector (every rank gets len // world rays)."""
    n = (indices.shape[0] // world) * world
    return indices[:n][rank::world]


def batch_indices(n_rays_total, step, rn_global, perm=None):
    """The reference's per-step slice of the shuffled ray table (ZT:451-453, wrap-around as in ZT:448-450)."""
    start = (step * rn_global) % max(n_rays_total - rn_global + 1, 1)
    idx = torch.arange(start, start + rn_global)
    return perm[idx] if perm is not None else idx


class FlatParameters:
    """All parameters and gradients of a module as views into two flat fp32 buffers: one all-reduce and one Adam
    launch per step.  Parameter objects (and their names) are untouched, so state_dict round-trips."""

    def __init__(self, module):
        ps = [p for p in module.parameters()]
        self.params = ps
        total = sum(p.numel() for p in ps)
        dev = ps[0].device
        self.flat = torch.zeros(total, device=dev)
        self.grad = torch.zeros(total, device=dev)
        off = 0
        for p in ps:
            n = p.numel()
            self.flat[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + n].view_as(p.data)
            p.grad = self.grad[off:off + n].view_as(p.data)
            off += n
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        self.t = 0

    def zero_grad(self):
        self.grad.zero_()

    def numel(self):
        return self.flat.numel()


def cuda_adam(fp: FlatParameters, lr, betas=(0.9, 0.999), eps=1e-8):
    """torch.optim.Adam semantics (train/lr_common_manager.py:11-15) in one launch of nunerf_adam."""
    from . import _lib
    if not fp.flat.is_cuda:
        raise RuntimeError("nu_nerf_b200.dist.cuda_adam needs CUDA parameters (no CPU fallback)")
    fp.t += 1
    _lib.call("nunerf_adam", fp.flat.data_ptr(), fp.grad.data_ptr(), fp.m.data_ptr(), fp.v.data_ptr(), fp.numel(),
              float(lr), betas[0], betas[1], eps, fp.t)


def warm_up_cos_lr(step, lr=5e-4, warm=5000, end=300000, alpha=0.05):
    """train/lr_common_manager.py:36-46 (WarmUpCosLR)."""
    if step < warm:
        return lr * step / warm
    prog = (step - warm) / (end - warm)
    return lr * ((math.cos(math.pi * prog) + 1.0) * 0.5 * (1 - alpha) + alpha)


def global_count(n_local, device, group=None):
    """Sum of a per-rank integer count over the group (one scalar all-reduce)."""
    # torch.full is a fill kernel: no pageable host->device copy (which would stall the host until the stream drains)
    t = torch.full((1,), float(n_local), device=device)
    _, world = world_info(group)
    if world > 1:
        dist.all_reduce(t, group=group)
    return t[0]


def init_sdf_reg(out, step, reg_step=1000, small_threshold=0.1, large_threshold=1.05):
    """InitSDFRegLoss (network/loss.py:111-140) on the renderer's `sdf_pts` / `sdf_vals` warm-up outputs (emitted while
    step < 1000): pushes the SDF below |x| - 0.1 near the origin and above |x| - 1.05 outside the unit sphere, annealed by
    (cos(pi step / 1000) + 1) / 2.  The reference's normalisations are reproduced as written (the `small` term divides
    its mean by (mean > 1e-5) + 1e-3).  Returns loss_sdf_large + loss_sdf_small (0-d), or None outside the warm-up."""
    if "sdf_vals" not in out or "sdf_pts" not in out or step >= reg_step:
        return None
    norm = torch.norm(out["sdf_pts"], dim=-1)
    sdf = out["sdf_vals"]
    # masked sums instead of boolean indexing: no host synchronisation (an empty mask gives 0, as in the reference)
    m_s = (norm < small_threshold).float()
    small = (torch.clamp(sdf - (norm - small_threshold), min=0.0) * m_s).sum() / m_s.sum().clamp_min(1.0)
    small = small / ((small > 1e-5).float() + 1e-3)
    m_l = (norm > large_threshold).float()
    large = torch.clamp((norm - large_threshold) - sdf, min=0.0) * m_l
    large = large.sum() / ((large > 1e-5).float().sum() + 1e-3)
    return (large + small) * ((math.cos(step / reg_step * math.pi) + 1.0) / 2.0)


def stage1_loss(out, loss_rgb, r_global, group=None, eikonal_weight=0.1, step=0, occ_loss_step=None,
                outer_reg_weight=0.5, share=1.0, n_in_global=None, outer_reg_step=15000, sdf_reg=True,
                normal_ori=False, mask_weight=0.01):
    """Trainer loss (trainer_zero.py:157-161 over the loss.py adapters of spherepot.yaml: nerf_render, eikonal, std,
    init_sdf_reg, occ, mask, outer_reg; normal_ori=True adds the normal-orientation term of the non-zero-thickness stage-1
    configs) with global denominators.
    `out` is the renderer's outputs dict of THIS rank (or of one chunk of its rays: `share` = the chunk's fraction of
    the rank's rays), `loss_rgb` [R_chunk]; returns the local share whose SUM over ranks (and chunks) is the global
    loss, so gradients are summed, not averaged, across ranks."""
    dev = loss_rgb.device
    _, world = world_info(group)
    loss = loss_rgb.sum() / r_global
    gerr = out["gradient_error"]
    has_inner = "transmission" in out            # the reference emits a zeros(1) placeholder when no sample is inside
    if n_in_global is not None:
        # the caller knows the number of inner samples of the WHOLE step (all chunks, all ranks): exact global mean
        if has_inner:
            loss = loss + eikonal_weight * gerr.sum() / torch.clamp(n_in_global, min=1.0)
    else:
        # every rank takes part in the count all-reduce, also the ones without inner samples
        n_in = global_count(gerr.shape[0] if has_inner else 0, dev, group)
        if has_inner:
            loss = loss + eikonal_weight * share * gerr.sum() / torch.clamp(n_in, min=1.0)
    if occ_loss_step is not None and step >= occ_loss_step and "loss_occ" in out:
        # OccLoss (loss.py:97-98) is a mean over the probed samples of one rank; ranks (and chunks) are averaged
        loss = loss + share * out["loss_occ"].mean() / world
    if step >= outer_reg_step and "color_bkgr" in out:
        # OuterRegLoss has its own hard-coded gate (loss.py:206: step >= 15000), independent of occ_loss_step
        sq = ((out["color_bkgr"] - out["color_spec"]) ** 2).sum()
        if "loss_normal" in out:
            # the stage-1 renderer of network/renderer.py reports both colours for the candidate rays only (NZ:798-821):
            # the mean runs over the candidates of all ranks (of this chunk: chunks are share-weighted)
            n_c = global_count(out["color_bkgr"].shape[0], dev, group)
            loss = loss + outer_reg_weight * share * sq / (3.0 * torch.clamp(n_c, min=1.0))
        else:
            loss = loss + outer_reg_weight * sq / (3.0 * r_global)
    if normal_ori and "loss_normal" in out:
        # NormalOrientationLoss (loss.py:101-112, configs/shape/real/ballstatue.yaml:17): mean over the rays
        loss = loss + out["loss_normal"].sum() / r_global
    if mask_weight and "loss_mask" in out:
        # MaskLoss (loss.py:152-164): the renderer's l1(masks, acc) is a mean over one rank's (chunk's) rays
        loss = loss + mask_weight * share * out["loss_mask"].reshape(()) / world
    if sdf_reg:
        # InitSDFRegLoss (first 1000 steps): its count-based denominators are per batch; ranks / chunks are averaged
        reg = init_sdf_reg(out, step)
        if reg is not None:
            loss = loss + share * reg / world
    return loss


def all_reduce_gradients(fp: FlatParameters, group=None):
    """SUM all-reduce of the flat gradient (the losses already carry the global denominators)."""
    _, world = world_info(group)
    if world > 1:
        dist.all_reduce(fp.grad, group=group)


class DataParallelTrainer:
    """One optimisation step of the reference trainer, ray-sharded over the ranks of `group`.

    render_fn(rays_o, rays_d, near, far, step) -> outputs dict   (NeROShapeRenderer.render on this rank's rays)
    rgb_loss_fn(pred, gt) -> [R_local]                            (NeROShapeRenderer.compute_rgb_loss)
    adam_fn(flat_params, lr)                                      (default: the CUDA Adam kernel)
    """

    def __init__(self, module, render_fn, rgb_loss_fn, adam_fn=cuda_adam, lr_fn=warm_up_cos_lr, group=None,
                 eikonal_weight=0.1, occ_loss_step=None, sample_fn=None, core_fn=None, count_fn=None, outer_reg_step=15000,
                 normal_ori=False, mask_weight=0.01):
        """sample_fn(rays_o, rays_d, near, far, step) -> z_vals; count_fn(rays_o, rays_d, z_vals) -> number of inner samples
        (a 0-d device tensor); core_fn(rays_o, rays_d, z_vals, step) -> outputs dict.  When the three are given and a step
        is split into several chunks, the step runs in two phases -- sample every chunk and count its inner samples
        (no gradient), all-reduce the total once, then render / differentiate chunk by chunk with the eikonal term divided
        by that GLOBAL count -- so that chunking does not change the loss (without them the eikonal mean of a chunked step
        is the ray-share weighted mean of per-chunk means)."""
        self.fp = FlatParameters(module)
        self.render_fn, self.rgb_loss_fn, self.adam_fn, self.lr_fn = render_fn, rgb_loss_fn, adam_fn, lr_fn
        self.sample_fn, self.core_fn, self.count_fn = sample_fn, core_fn, count_fn
        self.group = group
        self.eikonal_weight, self.occ_loss_step, self.outer_reg_step = eikonal_weight, occ_loss_step, outer_reg_step
        # loss terms of the non-zero-thickness stage-1 configs (loss list of configs/shape/real/ballstatue.yaml:17): the
        # normal-orientation term and, for renderers that emit `loss_mask`, the mask term (see stage1_loss)
        self.normal_ori, self.mask_weight = normal_ori, mask_weight
        self.last = {}

    def step(self, rays_o, rays_d, rgbs, near, far, step, chunk=None):
        rank, world = world_info(self.group)
        r_local = rays_o.shape[0]
        r_global = r_local * world
        chunk = r_local if chunk is None else min(chunk, r_local)
        self.fp.zero_grad()
        total = torch.zeros((), device=rays_o.device)
        slices = [slice(c0, min(r_local, c0 + chunk)) for c0 in range(0, r_local, chunk)]
        two_phase = len(slices) > 1 and self.sample_fn is not None and self.core_fn is not None and self.count_fn is not None
        zs, n_in_global = None, None
        if two_phase:
            with torch.no_grad():
                zs = [self.sample_fn(rays_o[sl], rays_d[sl], near[sl], far[sl], step) for sl in slices]
                n_loc = torch.stack([self.count_fn(rays_o[sl], rays_d[sl], z) for sl, z in zip(slices, zs)]).sum().float()
            n_in_global = n_loc.reshape(1).clone()
            if world > 1:
                dist.all_reduce(n_in_global, group=self.group)
            n_in_global = n_in_global[0]
        n_in_seen = 0
        for ci, sl in enumerate(slices):
            if two_phase:
                out = self.core_fn(rays_o[sl], rays_d[sl], zs[ci], step)
            else:
                out = self.render_fn(rays_o[sl], rays_d[sl], near[sl], far[sl], step)
            loss_rgb = self.rgb_loss_fn(out["ray_rgb"], rgbs[sl])
            share = (sl.stop - sl.start) / r_local
            loss = stage1_loss(out, loss_rgb, r_global, self.group, self.eikonal_weight, step, self.occ_loss_step,
                               share=share, n_in_global=n_in_global, outer_reg_step=self.outer_reg_step,
                               normal_ori=self.normal_ori, mask_weight=self.mask_weight)
            loss.backward()
            total = total + loss.detach()
            n_in_seen += int(out["gradient_error"].shape[0]) if "transmission" in out else 0
            self.last = {"n_in": int(out["gradient_error"].shape[0]) if "transmission" in out else 0}
        self.last["n_in_step"] = n_in_seen
        all_reduce_gradients(self.fp, self.group)
        self.adam_fn(self.fp, self.lr_fn(step))
        return total

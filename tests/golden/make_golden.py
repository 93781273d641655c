"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU through
oracle/ref_harness.py.  Run in the build container only:  python tests/golden/make_golden.py

Fixtures (all small):
  stage1_init.npz      per-tensor fingerprints of the reference state_dict after torch.manual_seed(0)
                       (the product must reproduce the reference initialisation exactly).
  stage1_train_R64.npz reference NeROShapeRenderer.render(..., step=10000, is_nerf=True) on the SURVEY 8(d)
                       synthetic rays: z_vals, every searchsorted index / sort permutation of sample_ray,
                       outputs dict, and strided samples of every parameter gradient of the trainer loss.
  stage1_sphere_R64.npz same with near/far from the unit sphere and perturb=0 (eval-style sampling).
  stage1_invs300_R64.npz the training case with deviation_network.variance set so that inv_s = 300 (a trained-like field:
                       the sdf -> alpha map is 15x sharper than at initialisation)
  fg_lut_reference.npz the reference's FG_LUT buffer (its asset assets/bsdf_256_256.bin as loaded by field.py:583)
  stage1_occ_R64.npz   same rays at step 20000: occlusion-probe loss (ZT:695-723) with its recorded randperm draw,
                       outer_reg, trainable inv_s.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
GRAD_STRIDE_SAMPLES = 64


def fingerprint(t):
    t = t.detach().double().reshape(-1)
    idx = torch.linspace(0, t.numel() - 1, min(t.numel(), 8)).long()
    return np.concatenate([[t.sum().item(), t.abs().sum().item(), (t * t).sum().item()], t[idx].numpy()])


def strided(t):
    t = t.detach().reshape(-1)
    idx = torch.linspace(0, t.numel() - 1, min(t.numel(), GRAD_STRIDE_SAMPLES)).long()
    return t[idx].numpy(), idx.numpy()


class Recorder:
    """Records torch.searchsorted / torch.sort results issued inside sample_ray (field.py:484, ZT:561)."""

    def __init__(self):
        self.inds, self.perms, self.sorted = [], [], []

    def __enter__(self):
        self._ss, self._sort = torch.searchsorted, torch.sort

        def ss(*a, **k):
            r = self._ss(*a, **k)
            self.inds.append(r.clone())
            return r

        def so(*a, **k):
            r = self._sort(*a, **k)
            self.sorted.append(r[0].clone())
            self.perms.append(r[1].clone())
            return r
        torch.searchsorted, torch.sort = ss, so
        return self

    def __exit__(self, *e):
        torch.searchsorted, torch.sort = self._ss, self._sort


class RandpermRecorder:
    """torch.randperm (ZT:710) replaced by a seeded draw that is recorded for the fixture."""

    def __init__(self, seed=4):
        self.g, self.perms = torch.Generator().manual_seed(seed), []

    def __enter__(self):
        self._orig = torch.randperm

        def rp(n, *a, **k):
            p = self._orig(n, generator=self.g)
            self.perms.append(p.clone())
            return p
        torch.randperm = rp
        return self

    def __exit__(self, *e):
        torch.randperm = self._orig


def run_case(net, R, sphere, perturb, step=10000):
    o, d = rh.synthetic_rays(R)
    U0, U1 = rh.synthetic_uniforms(R)
    gt = rh.synthetic_targets(R)
    if sphere:
        near, far = net.near_far_from_sphere(o, d)
    else:
        near, far = torch.full((R, 1), 0.8), torch.full((R, 1), 4.5)
    poses = torch.eye(3, 4)[None].repeat(R, 1, 1)
    net.zero_grad()
    # record the inputs of every upsample round as well
    ups_in = []
    orig_up = net.upsample

    def up(rays_o, rays_d, z_vals, sdf, n_imp, inv_s):
        ups_in.append((z_vals.clone(), sdf.clone().reshape(z_vals.shape), inv_s.reshape(-1)[0].clone()))
        return orig_up(rays_o, rays_d, z_vals, sdf, n_imp, inv_s)
    net.upsample = up
    with Recorder() as rec, rh.injected_rand([U0, U1] if perturb else []):
        z = net.sample_ray(o, d, near, far, 1.0 if perturb else 0.0)
    net.upsample = orig_up
    with RandpermRecorder() as rp:
        out = net.render_core(o, d, z, poses, cos_anneal_ratio=net.get_anneal_val(step), step=step, is_train=True,
                              is_nerf=True)
    loss = net.compute_rgb_loss(out["ray_rgb"], gt).mean() + (0.1 * out["gradient_error"]).mean()
    if step >= net.cfg["occ_loss_step"]:
        # trainer rule with the spherepot loss list (loss.py:97-98, :206-209)
        loss = loss + out["loss_occ"].mean() + 0.5 * torch.nn.functional.mse_loss(out["color_bkgr"].flatten(),
                                                                                  out["color_spec"].flatten())
    loss.backward()
    res = {"o": o, "d": d, "near": near, "far": far, "U0": U0, "U1": U1, "gt": gt, "z_vals": z,
           "loss": loss.detach(), "step": torch.tensor(step),
           "cos_anneal": torch.tensor(float(net.get_anneal_val(step)))}
    for i in range(4):
        res[f"inds_{i}"] = rec.inds[i].int()
        res[f"perm_{i}"] = rec.perms[i].int()
        res[f"z_merged_{i}"] = rec.sorted[i]
        res[f"z_in_{i}"], res[f"sdf_in_{i}"], res[f"inv_s_{i}"] = ups_in[i]
    for k in ["ray_rgb", "gradient_error", "acc", "color_bkgr", "color_spec", "std", "transmission", "metallic",
              "loss_occ"]:
        res["out_" + k] = out[k].detach()
    if rp.perms:
        res["occ_perm"] = rp.perms[0]
    res["occ_loss_max_pn"] = torch.tensor(net.cfg["occ_loss_max_pn"])
    for name, p in net.named_parameters():
        if p.grad is None:
            continue
        vals, idx = strided(p.grad)
        res["grad/" + name] = vals
        res["gradnorm/" + name] = np.array(p.grad.double().norm().item())
    return {k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in res.items()}


def main():
    # the reference loads ITS OWN split-sum table (assets/bsdf_256_256.bin, field.py:583); the table is committed as a
    # fixture so that the GPU box (no /root/reference) shades with exactly the values these goldens were produced with
    net, cfg = rh.load_stage1(seed=0)
    sd = net.state_dict()
    np.savez_compressed(os.path.join(OUT, "fg_lut_reference.npz"), FG_LUT=sd["color_network.FG_LUT"].numpy())
    np.savez_compressed(os.path.join(OUT, "stage1_init.npz"),
                        **{k: fingerprint(v) for k, v in sd.items() if k != "color_network.FG_LUT"})
    np.savez_compressed(os.path.join(OUT, "stage1_train_R64.npz"), **run_case(net, 64, sphere=False, perturb=True))
    np.savez_compressed(os.path.join(OUT, "stage1_sphere_R64.npz"), **run_case(net, 64, sphere=True, perturb=False))
    # step 20000: occlusion-probe loss + outer_reg active, inv_s trainable; occ_loss_max_pn lowered so that the
    # randperm sub-sampling branch (ZT:708-714) is exercised with 64 rays
    net.cfg["occ_loss_max_pn"] = 48
    np.savez_compressed(os.path.join(OUT, "stage1_occ_R64.npz"), **run_case(net, 64, sphere=False, perturb=True,
                                                                            step=20000))
    net.cfg["occ_loss_max_pn"] = 2048
    # a trained-like sharpness: inv_s = exp(10 variance) = 300 (ZT:657-685; random init has inv_s ~ 20), same rays
    import math
    v0 = net.deviation_network.variance.data.clone()
    net.deviation_network.variance.data.fill_(math.log(300.0) / 10.0)
    np.savez_compressed(os.path.join(OUT, "stage1_invs300_R64.npz"), **run_case(net, 64, sphere=False, perturb=True))
    net.deviation_network.variance.data.copy_(v0)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()

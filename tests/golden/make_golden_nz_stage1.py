"""Generates tests/golden/stage1nz_*.npz by running the UNMODIFIED stage-1 renderer of the non-zero-thickness module
(/root/reference/network/renderer.py:102-905) on CPU through oracle/ref_harness.py.
Run in the build container only:  python tests/golden/make_golden_nz_stage1.py

  stage1nz_R64.npz       NeROShapeRenderer.render_core (NZ:738-859) on the 64 synthetic rays / sample depths of the stage-1
                         fixtures (step 10000, is_nerf): ray_rgb, acc, gradient_error, loss_normal (NZ:766-780), color_bkgr /
                         color_spec on the candidate rays (NZ:798-821), transmission, metallic; the loss
                            mean charbonnier + 0.1 mean eikonal + mean loss_normal + 0.5 mse(color_bkgr, color_spec)
                            + 0.5 l1(masks, acc)                                (loss.py:105-112, :162-163, :206-209, NZ:478)
                         and strided samples + norms of every parameter gradient.
  stage1nz_sph_R64.npz   the same with shader_config.sphere_direction: true (144-wide outer light; the specular probe sees
                         [IDE(d) | IDE(exit direction)], NZ:805-809).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402
from make_golden import strided  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def case(sph):
    over = {"shader_config": {"sphere_direction": bool(sph), "human_light": False}}     # NZ:809 reads the key
    net, cfg = rh.load_stage1(seed=0, cfg_overrides=over, thick=True)
    R, step = 64, 10000
    o, d = rh.synthetic_rays(R)
    # the last 8 rays pass the unit sphere at a distance: their 65th sample is outside (not candidates of the specular probe)
    g8 = torch.Generator().manual_seed(21)
    d[-8:] = torch.nn.functional.normalize(-o[-8:] + 1.6 * torch.nn.functional.normalize(torch.randn(8, 3, generator=g8), dim=-1), dim=-1)
    U0, U1 = rh.synthetic_uniforms(R)
    gt = rh.synthetic_targets(R)
    masks = torch.rand(R, generator=torch.Generator().manual_seed(9))
    near, far = torch.full((R, 1), 0.8), torch.full((R, 1), 4.5)
    poses = torch.eye(3, 4)[None].repeat(R, 1, 1)
    net.zero_grad()
    with rh.injected_rand([U0, U1]):
        z = net.sample_ray(o, d, near, far, 1.0)
    out = net.render_core(o, d, z, poses, cos_anneal_ratio=net.get_anneal_val(step), step=step, is_train=True, is_nerf=True)
    loss = net.compute_rgb_loss(out["ray_rgb"], gt).mean() + (0.1 * out["gradient_error"]).mean() \
        + out["loss_normal"].mean() \
        + 0.5 * torch.nn.functional.mse_loss(out["color_bkgr"].flatten(), out["color_spec"].flatten()) \
        + 0.5 * torch.nn.functional.l1_loss(masks, out["acc"], reduction="mean")
    loss.backward()
    res = {"o": o, "d": d, "z_vals": z, "gt": gt, "masks": masks, "loss": loss.detach(), "step": torch.tensor(step),
           "cos_anneal": torch.tensor(float(net.get_anneal_val(step)))}
    for k in ["ray_rgb", "gradient_error", "loss_normal", "acc", "color_bkgr", "color_spec", "std", "transmission",
              "metallic"]:
        res["out_" + k] = out[k].detach()
    for name, p in net.named_parameters():
        if p.grad is None:
            continue
        vals, idx = strided(p.grad)
        res["grad/" + name] = vals
        res["gradnorm/" + name] = np.array(p.grad.double().norm().item())
    print("sphere_direction", sph, "loss", float(loss), "loss_normal mean", float(out["loss_normal"].mean()),
          "candidates", out["color_spec"].shape[0], "of", R)
    return {k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in res.items()}


def main():
    np.savez_compressed(os.path.join(OUT, "stage1nz_R64.npz"), **case(False))
    np.savez_compressed(os.path.join(OUT, "stage1nz_sph_R64.npz"), **case(True))


if __name__ == "__main__":
    main()

"""Stage-1 training step of the NON-zero-thickness module (nu_nerf_b200/renderer.NeROShapeRenderer; network/renderer.py:102-905)
at config-2 size (4096 rays, bf16 mode) next to the zero-thickness class on the same rays: what loss_normal (a second pass of
the compositing kernels, forward and backward), the candidate-ray probe and -- with --sph -- the sphere_direction shader
variant (192-column outer-light rows, variant encode kernels) cost.  Forward + backward of the trainer loss, no optimiser."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nu_nerf_b200 import synthetic as syn  # noqa: E402
from nu_nerf_b200.renderer import name2renderer as nz  # noqa: E402
from nu_nerf_b200.renderer_zerothick import name2renderer as zt, load_default_cfg  # noqa: E402


def step_ms(net, o, d, gt, masks, near, far, normal, iters=10, warm=3):
    def step():
        net.zero_grad(set_to_none=True)
        out = net.render(o, d, near, far, None, -1, 0.2, is_train=True, step=10000, is_nerf=True)
        loss = net.compute_rgb_loss(out["ray_rgb"], gt).mean() + (0.1 * out["gradient_error"]).mean()
        if normal:
            loss = loss + out["loss_normal"].mean() + 0.01 * torch.nn.functional.l1_loss(masks, out["acc"])
        loss.backward()
        return out
    for _ in range(warm):
        out = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    assert torch.isfinite(out["ray_rgb"]).all()
    return e0.elapsed_time(e1) / iters, out


def run(R=4096, sph=False):
    o, d = syn.synthetic_rays(R, seed=1)
    gt = syn.synthetic_targets(R, seed=3)
    o, d, gt = o.cuda(), d.cuda(), gt.cuda()
    masks = torch.rand(R, device="cuda")
    near, far = torch.full((R, 1), 0.8, device="cuda"), torch.full((R, 1), 4.5, device="cuda")
    res = {"workload": f"stage-1 render + trainer-loss backward, {R} rays, bf16 mode, sphere_direction={sph}"}
    for name, reg, normal in (("zero_thickness", zt, False), ("nonzero_thickness", nz, True)):
        if name == "zero_thickness" and sph:
            continue                                    # renderer_zerothick.py cannot run its stage 1 under the flag
        cfg = load_default_cfg()
        cfg["precision"] = "bf16"
        cfg["shader_config"] = {"sphere_direction": bool(sph), "human_light": False}
        torch.manual_seed(0)
        net = reg["shape"](cfg, training=False).cuda()
        ms, out = step_ms(net, o, d, gt, masks, near, far, normal)
        res[name] = {"ms_per_step": ms, "rays_per_s": R / ms * 1e3}
        if normal:
            res[name]["candidate_rays"] = int(out["color_spec"].shape[0])
            res[name]["loss_normal_mean"] = float(out["loss_normal"].detach().mean())
        del net
        torch.cuda.empty_cache()
    return res


if __name__ == "__main__":
    print(json.dumps(run(4096, sph="--sph" in sys.argv)))

"""Helper of tests/test_reference_adapters_gpu.py, run as a SUBPROCESS (oracle/ref_harness.install() monkey-patches torch
for the whole process): renders with the product on cuda:0, then feeds the outputs dicts to the UNMODIFIED reference's
network/loss.py adapters and network/metrics.py (from oracle/_ref -- it travels to the GPU box -- or /root/reference)
and prints one JSON line."""
import importlib.util
import json
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
DEV = "cuda"


def ref_root():
    for c in (os.environ.get("NUNERF_REFERENCE_ROOT"), os.path.join(ROOT, "oracle", "_ref"), "/root/reference"):
        if c and os.path.isdir(os.path.join(c, "network")):
            return c
    return None


def main():
    root = ref_root()
    if root is None:
        print(json.dumps({"skip": "no reference tree (oracle/_ref is built by __graft_entry__.build() in the container)"}))
        return
    from nu_nerf_b200 import dist, feeder
    from nu_nerf_b200.renderer_zerothick import NeROShapeRenderer, load_default_cfg
    from nu_nerf_b200.synthetic import make_stage2, synthetic_rays, synthetic_targets
    spec = importlib.util.spec_from_file_location("ref_loss", os.path.join(root, "network", "loss.py"))
    ref_loss = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_loss)
    res = {"root": root, "train": {}}
    # ---------------- stage 1: trainer loss through the reference adapters of configs/shape/nerf/spherepot.yaml:13
    torch.manual_seed(0)
    cfg = load_default_cfg()
    cfg["precision"] = "bf16"
    net = NeROShapeRenderer(cfg, training=False).cuda()
    R = 256
    o, d = (t.to(DEV) for t in synthetic_rays(R))
    gt = synthetic_targets(R).to(DEV)
    near, far = torch.full((R, 1), 0.8, device=DEV), torch.full((R, 1), 4.5, device=DEV)
    names = ["nerf_render", "eikonal", "std", "init_sdf_reg", "occ", "mask", "outer_reg"]
    for step in (500, 10000, 20000):
        out = net.render(o, d, near, far, None, -1, net.get_anneal_val(step), is_train=True, step=step, is_nerf=True)
        out["loss_rgb"] = net.compute_rgb_loss(out["ray_rgb"], gt)              # added by train_step (ZT:468)
        log = {}
        for n in names:
            log.update(ref_loss.name2loss[n]({})(out, None, step))
        total = sum(torch.mean(v) for k, v in log.items() if k.startswith("loss"))   # trainer_zero.py:157-161
        ours = dist.stage1_loss(out, out["loss_rgb"], R, eikonal_weight=0.1, step=step, occ_loss_step=cfg["occ_loss_step"])
        # the remaining adapters read keys the renderer emits, too
        extra = {}
        for n in ("transmission_reg", "metallic_reg", "mat_reg", "normal_ori"):
            extra.update(ref_loss.name2loss[n]({})(out, None, step))
        res["train"][str(step)] = {"reference_total": float(total), "ours": float(ours), "keys": sorted(log),
                                   "extra_keys": sorted(extra), "shapes": {k: list(v.shape) for k, v in log.items()}}
    # ---------------- stage 2: configs/stage2/nerf/spherepot.yaml:16
    net2 = make_stage2("bf16").cuda()
    o2, d2 = o[:64].contiguous(), d[:64].contiguous()
    out2 = net2.render(o2, d2, None, None, None, -1, 0.2, is_train=True, step=10000, is_nerf=True)
    tm = out2["tir_mask"]
    out2["loss_rgb"] = net2.compute_rgb_loss(out2["ray_rgb"] * tm, gt[:64] * tm)
    log = {}
    for n in ("eikonal", "std", "nerf_render"):
        log.update(ref_loss.name2loss[n]({"eikonal_weight": 0.02})(out2, None, 10000))
    total2 = sum(torch.mean(v) for k, v in log.items() if k.startswith("loss"))
    ours2 = out2["loss_rgb"].mean() + (0.02 * out2["gradient_error"]).mean()
    res["stage2_train"] = {"reference_total": float(total2), "ours": float(ours2), "keys": sorted(log)}
    # ---------------- eval outputs -> network/metrics.py
    h, w = 6, 10
    g = torch.Generator().manual_seed(2)
    imgs = torch.rand(1, 3, h, w, generator=g).to(DEV)
    K = torch.tensor([[20.0, 0, 5.0], [0, 20.0, 3.0], [0, 0, 1]])[None].to(DEV)
    c2w = torch.eye(3, 4)[None].clone()
    c2w[0, 2, 3] = 3.0
    net.set_eval_source(feeder.image_eval_source(imgs, K, c2w.to(DEV), is_nerf=True))
    net.is_nerf = True
    ev1 = {k: v.detach().cpu() for k, v in net({"eval": True, "index": 0, "step": 20000}).items() if torch.is_tensor(v)}
    net2.set_eval_source(feeder.image_eval_source(imgs, K, c2w.to(DEV), is_nerf=True))
    ev2 = {k: v.detach().cpu() for k, v in net2({"eval": True, "index": 0, "step": 10000}).items() if torch.is_tensor(v)}
    os.environ["NUNERF_REFERENCE_ROOT"] = root
    from oracle import ref_harness as rh
    rh.install()
    from network.metrics import ShapeRenderMetrics, Stage2RenderMetrics, draw_materials, draw_materials_s2
    os.chdir(tempfile.mkdtemp())
    m1 = ShapeRenderMetrics({})(ev1, None, 20000, data_index=0, model_name="adapters")
    m2 = Stage2RenderMetrics({})(ev2, None, 10000, data_index=0, model_name="adapters")
    mat_keys = ['diffuse_albedo', 'diffuse_light', 'diffuse_color', 'refraction_light', 'specular_albedo', 'specular_light',
                'specular_color', 'specular_ref', 'transmission_weight', 'roughness', 'occ_prob', 'indirect_light']
    res["metrics"] = {"psnr1": float(m1["psnr"][0]), "psnr2": float(m2["psnr"][0]),
                      "material_keys_present": [k for k in mat_keys if k in ev1],
                      "s2_keys_present": [k for k in ("specular_light", "specular_color", "specular_ref") if k in ev2],
                      "panels1": len(draw_materials(ev1, h, w)), "panels2": len(draw_materials_s2(ev2, h, w))}
    print(json.dumps(res))


if __name__ == "__main__":
    main()

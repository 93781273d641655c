"""Generates tests/golden/stage2nz_*.npz by running the UNMODIFIED non-zero-thickness reference Stage2Renderer
(/root/reference/network/renderer.py:907-2378) on CPU through oracle/ref_harness.py (thick=True).
Run in the build container only:  python tests/golden/make_golden_nz.py

  stage2nz_init.npz        per-tensor fingerprints of the reference state_dict (seed 5, NZ stage-1 checkpoint from seed 0).
  stage2nz_sphere_R64.npz  outer mesh = UV sphere r 0.6 (positive curvature everywhere): ray_trace (NZ:1610-2148) per bounce --
                           the INPUTS of the shell-offset geometry as the reference saw them (hit point, interpolated normal,
                           interpolated Gaussian curvature, IoR / thickness network outputs, incoming direction) and its
                           outputs (pass masks, next origin / direction, IoR ratios, mesh normals, TIR mask, every sampled
                           path point) -- plus the render_core outputs dict (NZ:2155-2353) in train and eval mode.
  stage2nz_grads_R64.npz   sphere case with autograd on: trainer loss (mean charbonnier with the TIR mask + 0.02 * eikonal) and
                           strided samples + norms of every parameter gradient, incl. IORs_pred and thickness_pred.
  stage2nz_sph_*.npz       the three sphere fixtures again with shader_config.sphere_direction: true in both stages (144-wide
                           outer light, field.py:594-597, :641-651, :675-680) -- the shader variant of configs/*/real/*.yaml.
  stage2nz_torus_R96.npz   the same ray_trace record on a torus (both curvature signs, re-entering rays).

Vertex Gaussian curvature: oracle/ref_harness.angle_defect_curvature (the stated replacement for the PyMesh attribute;
parity against PyMesh itself unpinned).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402
from make_golden import fingerprint, strided  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def trace_record(net, o, d):
    """ray_trace with the per-bounce inputs of the shell geometry recorded from the reference's own calls."""
    hits, iors, thick = [], [], []
    orig_d = net.scene.Dintersect
    orig_i, orig_t = net.IORs_pred.forward, net.thickness_pred.forward

    def rec_d(ray):
        info, conv = orig_d(ray)
        hits.append(dict(o=ray.origin.clone(), d=ray.direction.clone(), conv=conv.clone(),
                         x=info.intersection_point.clone(), n=info.n.clone(), g_k=info.g_k.clone(),
                         tri=info.faces_ind.clone()))
        return info, conv

    def rec_i(x):
        y = orig_i(x)
        iors.append(y.clone())
        return y

    def rec_t(x):
        y = orig_t(x)
        thick.append(y.clone())
        return y
    net.scene.Dintersect = rec_d
    net.IORs_pred.forward, net.thickness_pred.forward = rec_i, rec_t
    try:
        with torch.no_grad():
            out = net.ray_trace(o, d, None)
    finally:
        net.scene.Dintersect = orig_d
        net.IORs_pred.forward, net.thickness_pred.forward = orig_i, orig_t
    pathes, converges, directions, ior_ratios, infinity_bkgr, gradient_mesh, tir_mask = out
    res = {"o": o.numpy(), "d": d.numpy(), "n_segments": np.array(len(pathes)), "n_bounces": np.array(len(hits)),
           "tir_mask": tir_mask.numpy()}
    for k in range(len(pathes)):
        res[f"path_{k}"] = pathes[k].numpy()
        res[f"converge_{k}"] = converges[k].numpy()
        res[f"bkgr_{k}"] = infinity_bkgr[k].numpy()
    for k in range(len(directions)):
        res[f"dir_{k}"] = directions[k].numpy()
    for k in range(len(ior_ratios)):
        res[f"ior_{k}"] = ior_ratios[k].numpy()
        res[f"nmesh_{k}"] = gradient_mesh[k].numpy()
    for k, h in enumerate(hits):
        res[f"in_o_{k}"], res[f"in_d_{k}"] = h["o"].numpy(), h["d"].numpy()
        res[f"in_hit_{k}"], res[f"in_tri_{k}"] = h["conv"].numpy(), h["tri"].numpy()
        res[f"in_x_{k}"], res[f"in_n_{k}"], res[f"in_gk_{k}"] = h["x"].numpy(), h["n"].numpy(), h["g_k"].numpy()
        res[f"in_ior_{k}"], res[f"in_thick_{k}"] = iors[k].numpy(), thick[k].numpy()
    return res, out


def sphere_case(tag, sphere_direction):
    V, Fc = rh.uv_sphere(radius=0.6, nu=48, nv=24)
    net, cfg = rh.load_stage2(V, Fc, thick=True, sphere_direction=sphere_direction)
    sd = net.state_dict()
    np.savez_compressed(os.path.join(OUT, f"stage2nz{tag}_init.npz"),
                        **{k: fingerprint(v) for k, v in sd.items() if not k.endswith("FG_LUT")})
    o, d = rh.synthetic_rays(64)
    res, out = trace_record(net, o, d)
    res.update(mesh_radius=np.array(0.6), mesh_nu=np.array(48), mesh_nv=np.array(24), step=np.array(10000),
               cos_anneal=np.array(0.2))
    pathes, converges, directions, ior_ratios, infinity_bkgr, gradient_mesh, tir_mask = out
    with torch.no_grad():
        for mode, is_train in (("train", True), ("eval", False)):
            r = net.render_core(o, d, pathes, converges, directions, infinity_bkgr, gradient_mesh, ior_ratios, None,
                                cos_anneal_ratio=0.2, step=10000, is_train=is_train, is_nerf=True)
            for kk, v in r.items():
                res[f"{mode}_{kk}"] = v.detach().float().numpy()
    if not sphere_direction:
        # step 20000: the occlusion-probe loss of the INNER field (NZ:2222-2230, :1580-1608); occ_sdf_thresh widened so that
        # a random-init field has candidates (|sdf| < 0.01 selects almost nothing there), no randperm sub-sampling
        net.cfg["occ_sdf_thresh"], net.cfg["occ_loss_max_pn"] = 0.05, 1 << 20
        net.zero_grad()
        r20 = net.render_core(o, d, pathes, converges, directions, infinity_bkgr, gradient_mesh, ior_ratios, None,
                              cos_anneal_ratio=0.2, step=20000, is_train=True, is_nerf=True)
        res["occ20_loss_occ"] = r20["loss_occ"].detach().reshape(-1).numpy()
        res["occ20_ray_rgb"] = r20["ray_rgb"].detach().numpy()
        r20["loss_occ"].mean().backward()
        n_g = 0
        for name, p_ in net.named_parameters():
            if p_.grad is not None and float(p_.grad.abs().sum()) > 0:
                vals, idx = strided(p_.grad)
                res["occ20_grad/" + name], res["occ20_gradnorm/" + name] = vals, np.array(p_.grad.double().norm().item())
                n_g += 1
        print("inner occlusion loss at step 20000:", res["occ20_loss_occ"], "tensors with a gradient:", n_g)
        net.cfg["occ_sdf_thresh"], net.cfg["occ_loss_max_pn"] = 0.01, 2048
    np.savez_compressed(os.path.join(OUT, f"stage2nz{tag}_sphere_R64.npz"), **res)
    # ---- gradients of the stage-2 trainer loss through ray_trace + render_core (autograd of the reference): the IoR AND
    # the thickness network receive theirs through the shell geometry
    gt = rh.synthetic_targets(64)
    net.zero_grad()
    pathes, converges, directions, ior_ratios, infinity_bkgr, gradient_mesh, tir_mask = net.ray_trace(o, d, None)
    r = net.render_core(o, d, pathes, converges, directions, infinity_bkgr, gradient_mesh, ior_ratios, None,
                        cos_anneal_ratio=0.2, step=10000, is_train=True, is_nerf=True)
    tm = tir_mask.detach()
    loss = net.compute_rgb_loss(r["ray_rgb"] * tm, gt * tm).mean() + (0.02 * r["gradient_error"]).mean()
    loss.backward()
    gres = {"gt": gt.numpy(), "loss": loss.detach().numpy(), "ray_rgb": r["ray_rgb"].detach().numpy()}
    for name, p_ in net.named_parameters():
        if p_.grad is None:
            continue
        vals, idx = strided(p_.grad)
        gres["grad/" + name] = vals
        gres["gradnorm/" + name] = np.array(p_.grad.double().norm().item())
    np.savez_compressed(os.path.join(OUT, f"stage2nz{tag}_grads_R64.npz"), **gres)
    print("grad norms:", {k: float(v) for k, v in gres.items() if k.startswith("gradnorm/") and
                          ("IORs_pred" in k or "thickness" in k) and k.endswith("bias")})
    print("sphere:", [p.shape for p in pathes], [int(c.sum()) for c in converges], int(tir_mask.sum()),
          "rgb", res["train_ray_rgb"].mean(0))



def zero_thickness_sphere_direction():
    """stage2_sph_R64.npz: the ZERO-thickness stage 2 (renderer_zerothick.py) with shader_config.sphere_direction: true in both
    stages (configs/stage2/real/eikonal_wineglass.yaml): ray_trace lists, render_core outputs, trainer-loss gradients."""
    V, Fc = rh.uv_sphere(radius=0.6, nu=48, nv=24)
    net, cfg = rh.load_stage2(V, Fc, thick=False, sphere_direction=True)
    o, d = rh.synthetic_rays(64)
    gt = rh.synthetic_targets(64)
    net.zero_grad()
    pathes, converges, directions, ior_ratios, infinity_bkgr, gradient_mesh, tir_mask = net.ray_trace(o, d)
    r = net.render_core(o, d, pathes, converges, directions, infinity_bkgr, gradient_mesh, ior_ratios, None,
                        cos_anneal_ratio=0.2, step=10000, is_train=True, is_nerf=True)
    tm = tir_mask.detach()
    loss = net.compute_rgb_loss(r["ray_rgb"] * tm, gt * tm).mean() + (0.02 * r["gradient_error"]).mean()
    loss.backward()
    res = {"o": o.numpy(), "d": d.numpy(), "gt": gt.numpy(), "n_segments": np.array(len(pathes)),
           "tir_mask": tir_mask.numpy(), "loss": loss.detach().numpy(), "train_ray_rgb": r["ray_rgb"].detach().numpy()}
    for k in range(len(pathes)):
        res[f"path_{k}"], res[f"converge_{k}"] = pathes[k].detach().numpy(), converges[k].numpy()
        res[f"bkgr_{k}"] = infinity_bkgr[k].numpy()
    for k in range(len(directions)):
        res[f"dir_{k}"] = directions[k].detach().numpy()
    for k in range(len(ior_ratios)):
        res[f"ior_{k}"], res[f"nmesh_{k}"] = ior_ratios[k].detach().numpy(), gradient_mesh[k].detach().numpy()
    for name, p_ in net.named_parameters():
        if p_.grad is None or name.startswith("IORs_pred"):
            continue
        vals, idx = strided(p_.grad)
        res["grad/" + name] = vals
        res["gradnorm/" + name] = np.array(p_.grad.double().norm().item())
    np.savez_compressed(os.path.join(OUT, "stage2_sph_R64.npz"), **res)
    print("zero-thickness + sphere_direction: loss", float(loss), "rgb", res["train_ray_rgb"].mean(0))


def main():
    zero_thickness_sphere_direction()
    sphere_case("", False)
    sphere_case("_sph", True)          # shader_config.sphere_direction: true (the real-data configs of this renderer)
    Vt, Ft = rh.torus()
    net2, _ = rh.load_stage2(Vt, Ft, thick=True)
    o, d = rh.synthetic_rays(96, seed=7)
    res2, out2 = trace_record(net2, o, d)
    res2.update(torus_R=np.array(0.55), torus_r=np.array(0.22), torus_nu=np.array(40), torus_nv=np.array(20))
    np.savez_compressed(os.path.join(OUT, "stage2nz_torus_R96.npz"), **res2)
    # gradients on the torus: the vertex curvatures differ across a triangle there, so the curvature radius of the shell is
    # itself a function of the refracted path (DiffRender.py:113-116) -- the IoR / thickness gradients contain that term
    gt = rh.synthetic_targets(96)
    net2.zero_grad()
    pathes, converges, directions, ior_ratios, infinity_bkgr, gradient_mesh, tir_mask = net2.ray_trace(o, d, None)
    r = net2.render_core(o, d, pathes, converges, directions, infinity_bkgr, gradient_mesh, ior_ratios, None,
                         cos_anneal_ratio=0.2, step=10000, is_train=True, is_nerf=True)
    tm = tir_mask.detach()
    loss = net2.compute_rgb_loss(r["ray_rgb"] * tm, gt * tm).mean() + (0.02 * r["gradient_error"]).mean()
    loss.backward()
    gres = {"gt": gt.numpy(), "loss": loss.detach().numpy(), "ray_rgb": r["ray_rgb"].detach().numpy()}
    for name, p_ in net2.named_parameters():
        if p_.grad is None:
            continue
        vals, idx = strided(p_.grad)
        gres["grad/" + name] = vals
        gres["gradnorm/" + name] = np.array(p_.grad.double().norm().item())
    np.savez_compressed(os.path.join(OUT, "stage2nz_torus_grads_R96.npz"), **gres)
    print("torus:", [p.shape for p in out2[0]], [int(c.sum()) for c in out2[1]], int(out2[6].sum()),
          "curvature signs", [(float((res2[f"in_gk_{k}"] >= 0).mean()) if res2[f"in_gk_{k}"].size else None)
                              for k in range(int(res2["n_bounces"]))])


if __name__ == "__main__":
    main()

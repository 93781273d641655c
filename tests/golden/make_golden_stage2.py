"""Generates tests/golden/stage2_*.npz by running the UNMODIFIED reference Stage2Renderer (zero-thickness,
/root/reference/network/renderer_zerothick.py:868-2011) on CPU through oracle/ref_harness.py.
Run in the build container only:  python tests/golden/make_golden_stage2.py

  stage2_init.npz   per-tensor fingerprints of the reference Stage2Renderer state_dict (seed 5, stage-1 checkpoint from
                    seed 0): the product must reproduce the reference initialisation exactly.
  stage2_R64.npz    SURVEY 8(d) config-4 style case on 64 rays: outer mesh = UV sphere r 0.6 (48 x 24, 2208
                    triangles), random-init nested inner field; ray_trace intermediates per bounce (hit masks, hit
                    triangle ids from the brute-force oracle, refracted directions, IoR ratios, mesh normals, TIR mask,
                    every sampled path point) and the outputs dict in train and eval mode.
  stage2_grads_R64.npz  same case with autograd on: trainer loss of configs/stage2/nerf/spherepot.yaml
                    (mean(loss_rgb with the TIR mask, ZT:1272) + mean(0.02 * gradient_error)) and strided samples + norms of
                    every parameter gradient (incl. IORs_pred, whose gradient flows through the path geometry).
  stage2_geomgrads_R64.npz  d loss / d (path points, segment directions, mesh normals, IoR ratios) of the same backward
                    pass (the reference's autograd; intermediate fixtures for the position-gradient kernels).
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
MESH = dict(radius=0.6, nu=48, nv=24)


def main():
    V, Fc = rh.uv_sphere(**MESH)
    net, cfg = rh.load_stage2(V, Fc)          # with the reference's own FG_LUT asset (fg_lut_reference.npz)
    sd = net.state_dict()
    np.savez_compressed(os.path.join(OUT, "stage2_init.npz"),
                        **{k: fingerprint(v) for k, v in sd.items() if not k.endswith("FG_LUT")})
    R = 64
    o, d = rh.synthetic_rays(R)
    res = {"o": o.numpy(), "d": d.numpy(), "mesh_radius": np.array(MESH["radius"]), "mesh_nu": np.array(MESH["nu"]),
           "mesh_nv": np.array(MESH["nv"]), "step": np.array(10000), "cos_anneal": np.array(0.2)}
    # hit triangle ids per bounce (the oracle's closest-hit rule), recorded from the shimmed optix_mesh.intersect
    hits = []
    orig = net.scene.optix_mesh.intersect

    def rec(ray):
        h, i = orig(ray)
        hits.append((ray.clone(), h.clone(), i.clone()))
        return h, i
    net.scene.optix_mesh.intersect = rec
    with torch.no_grad():
        pathes, converges, directions, ior_ratios, infinity_bkgr, gradient_mesh, tir_mask = net.ray_trace(o, d)
    n_seg = len(pathes)
    res["n_segments"] = np.array(n_seg)
    for k in range(n_seg):
        res[f"path_{k}"] = pathes[k].numpy()
        res[f"converge_{k}"] = converges[k].numpy()
        res[f"bkgr_{k}"] = infinity_bkgr[k].numpy()
        res[f"dir_{k}"] = directions[k].numpy()
        res[f"trace_ray_{k}"] = hits[k][0].numpy()
        res[f"trace_hit_{k}"] = hits[k][1].numpy()
        res[f"trace_tri_{k}"] = hits[k][2].numpy()
    res[f"dir_{n_seg}"] = directions[n_seg].numpy() if len(directions) > n_seg else np.zeros((0, 3), np.float32)
    for k in range(len(ior_ratios)):
        res[f"ior_{k}"] = ior_ratios[k].numpy()
        res[f"nmesh_{k}"] = gradient_mesh[k].numpy()
    res["tir_mask"] = tir_mask.numpy()
    with torch.no_grad():
        for mode, is_train in (("train", True), ("eval", False)):
            out = net.render_core(o, d, pathes, converges, directions, infinity_bkgr, gradient_mesh, ior_ratios, None,
                                  cos_anneal_ratio=0.2, step=10000, is_train=is_train, is_nerf=True)
            for kk, v in out.items():
                res[f"{mode}_{kk}"] = v.detach().float().numpy()
    np.savez_compressed(os.path.join(OUT, "stage2_R64.npz"), **res)
    # ---- gradients of the stage-2 trainer loss (autograd through ray_trace + render_core)
    gt = rh.synthetic_targets(R)
    net.zero_grad()
    pathes, converges, directions, ior_ratios, infinity_bkgr, gradient_mesh, tir_mask = net.ray_trace(o, d)
    # gradients w.r.t. the path geometry (what carries the loss to IORs_pred): keep them for the fixture
    geo = {}
    for k, t in enumerate(pathes):
        if t.requires_grad:
            t.retain_grad(); geo[f"d_path_{k}"] = t
    for k, t in enumerate(directions):
        if t.requires_grad:
            t.retain_grad(); geo[f"d_dir_{k}"] = t
    for k, t in enumerate(gradient_mesh):
        if t.requires_grad:
            t.retain_grad(); geo[f"d_nmesh_{k}"] = t
    for k, t in enumerate(ior_ratios):
        if t.requires_grad:
            t.retain_grad(); geo[f"d_ior_{k}"] = t
    out = net.render_core(o, d, pathes, converges, directions, infinity_bkgr, gradient_mesh, ior_ratios, None,
                          cos_anneal_ratio=0.2, step=10000, is_train=True, is_nerf=True)
    tm = tir_mask.detach()
    loss_rgb = net.compute_rgb_loss(out["ray_rgb"] * tm, gt * tm)
    loss = loss_rgb.mean() + (0.02 * out["gradient_error"]).mean()
    loss.backward()
    gres = {"gt": gt.numpy(), "loss": loss.detach().numpy(), "ray_rgb": out["ray_rgb"].detach().numpy()}
    ggeo = {k: (t.grad if t.grad is not None else torch.zeros_like(t)).detach().numpy() for k, t in geo.items()}
    np.savez_compressed(os.path.join(OUT, "stage2_geomgrads_R64.npz"), **ggeo)
    print("geometry gradients:", {k: (v.shape, float(np.abs(v).max()) if v.size else 0.0) for k, v in ggeo.items()})
    for name, p_ in net.named_parameters():
        if p_.grad is None:
            continue
        vals, idx = strided(p_.grad)
        gres["grad/" + name] = vals
        gres["gradnorm/" + name] = np.array(p_.grad.double().norm().item())
    # ---- conditioning of every parameter gradient: the reference's OWN gradient under operand rounding at the precision
    # of the product's fp32-accurate mode (bf16 hi + lo planes = 16 mantissa bits: relative rounding <= 2^-17).  Two
    # trials with every weight multiplied by 1 + 2^-17 u, u ~ U(-1, 1); recorded with the metrics of the parity test.
    base = {n_: p_.grad.detach().clone() for n_, p_ in net.named_parameters() if p_.grad is not None}
    saved = {n_: p_.detach().clone() for n_, p_ in net.named_parameters()}
    floors = {n_: [0.0, 0.0] for n_ in base}
    gen = torch.Generator().manual_seed(11)
    for trial in range(2):
        with torch.no_grad():
            for n_, p_ in net.named_parameters():
                if p_.dtype.is_floating_point and p_.numel() > 1:
                    p_.copy_(saved[n_] * (1.0 + 2.0 ** -17 * (2.0 * torch.rand(p_.shape, generator=gen) - 1.0)))
        net.zero_grad()
        lists = net.ray_trace(o, d)
        # same discrete trace as the unperturbed run is required for a like-for-like comparison: re-run ray_trace (the
        # geometry depends on IORs_pred) and keep the trial only if the hit / TIR masks are unchanged
        p2, c2, d2, i2, b2, g2, t2 = lists
        same = len(p2) == len(pathes) and all(torch.equal(a_, b_) for a_, b_ in zip(c2, converges)) and torch.equal(t2, tir_mask)
        if not same:
            print("trial", trial, "changed the discrete trace: skipped")
            continue
        out_t = net.render_core(o, d, p2, c2, d2, b2, g2, i2, None, cos_anneal_ratio=0.2, step=10000, is_train=True,
                                is_nerf=True)
        tm_t = t2.detach()
        (net.compute_rgb_loss(out_t["ray_rgb"] * tm_t, gt * tm_t).mean() + (0.02 * out_t["gradient_error"]).mean()).backward()
        for n_, p_ in net.named_parameters():
            if n_ not in base or p_.grad is None:
                continue
            g0, g1 = base[n_].reshape(-1), p_.grad.detach().reshape(-1)
            idx = torch.linspace(0, g0.numel() - 1, min(g0.numel(), 64)).long()
            nrm = g0.double().norm().item()
            if nrm == 0.0:
                continue
            scale = max(g0[idx].abs().max().item(), nrm / max(g0.numel(), 1) ** 0.5)
            floors[n_][0] = max(floors[n_][0], (g1[idx] - g0[idx]).abs().max().item() / scale)
            floors[n_][1] = max(floors[n_][1], abs(g1.double().norm().item() - nrm) / nrm)
    with torch.no_grad():
        for n_, p_ in net.named_parameters():
            p_.copy_(saved[n_])
    for n_, (f_s, f_n) in floors.items():
        gres["floor/" + n_] = np.array([f_s, f_n])
    worst = sorted(floors.items(), key=lambda kv: -kv[1][0])[:8]
    print("largest operand-rounding floors (sampled, norm):", [(k_, round(v_[0], 5), round(v_[1], 5)) for k_, v_ in worst])
    np.savez_compressed(os.path.join(OUT, "stage2_grads_R64.npz"), **gres)
    print("params with grad:", sum(1 for k in gres if k.startswith("grad/")), "loss", float(loss))
    for f in ("stage2_init.npz", "stage2_R64.npz", "stage2_grads_R64.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)))
    print({k: (v.shape, float(np.mean(v))) for k, v in res.items() if k.startswith(("train_", "eval_", "converge", "tir"))})


if __name__ == "__main__":
    main()

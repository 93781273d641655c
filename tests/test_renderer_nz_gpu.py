"""GPU parity of the non-zero-thickness Stage2Renderer (nu_nerf_b200/renderer.py; network/renderer.py:907-2378) against
outputs of the UNMODIFIED reference run through oracle/ref_harness.py (tests/golden/stage2nz_*.npz, made by
tests/golden/make_golden_nz.py; vertex curvature = the stated angle-defect definition on both sides).

Bars: hit masks, hit triangle ids, pass masks and the TIR mask bit-exact; interpolated curvature 1e-4 relative; directions /
IoR ratios / mesh normals 1e-5; path points 1e-4 on the uniformly sampled segments and the quantile gate of the CDF
inversion on the up-sampled one; rendered colour 1e-4 in the fp32-accurate mode when render_core (with the PE-8 / PE-2
AppShadingNetwork_SpecInner kernels) is fed the reference's own lists, 2e-3 end to end; parameter gradients of the trainer
loss -- including IORs_pred and thickness_pred, which receive theirs through the shell geometry -- against the reference's
autograd on equal sample parameters."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, make_stage2
from nu_nerf_b200.synthetic import torus

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(GOLDEN, "stage2nz_sphere_R64.npz"))


@pytest.fixture(scope="module")
def net():
    return make_stage2("split", thick=True).cuda()


def _lists(G):
    T = lambda k: torch.from_numpy(G[k]).to(DEV)
    n = int(G["n_segments"])
    return ([T(f"path_{k}") for k in range(n)], [T(f"converge_{k}") for k in range(n)],
            [T(f"dir_{k}") for k in range(n + 1)], [T(f"ior_{k}") for k in range(n) if f"ior_{k}" in G.files],
            [T(f"bkgr_{k}") for k in range(n)], [T(f"nmesh_{k}") for k in range(n) if f"nmesh_{k}" in G.files])


def _check_trace(net_, G):
    o, d = torch.from_numpy(G["o"]).to(DEV), torch.from_numpy(G["d"]).to(DEV)
    tr = {}
    with torch.no_grad():
        pathes, converges, directions, iors, bkgr, nmesh, tir = net_.ray_trace(o, d, None, trace=tr)
    n = int(G["n_segments"])
    assert len(pathes) == n and len(converges) == n and len(directions) == n + 1
    for k in range(int(G["n_bounces"])):
        hit_ref = torch.from_numpy(G[f"in_hit_{k}"]).float().flatten()
        assert torch.equal(tr[f"trace_hit_{k}"].cpu().float().flatten(), hit_ref), k
        got_tri = tr[f"trace_tri_{k}"].cpu().int().flatten()[hit_ref.bool()]
        assert torch.equal(got_tri, torch.from_numpy(G[f"in_tri_{k}"]).int().flatten()), k
    for k in range(n):
        assert torch.equal(converges[k].cpu(), torch.from_numpy(G[f"converge_{k}"])), k
        assert torch.equal(bkgr[k].cpu(), torch.from_numpy(G[f"bkgr_{k}"])), k
    for k in range(n + 1):
        ref = torch.from_numpy(G[f"dir_{k}"])
        assert directions[k].shape == ref.shape and (ref.numel() == 0 or (directions[k].cpu() - ref).abs().max().item() < 1e-5), k
    for k in range(len(iors)):
        assert (iors[k].cpu() - torch.from_numpy(G[f"ior_{k}"])).abs().max().item() < 1e-5, k
        assert (nmesh[k].cpu() - torch.from_numpy(G[f"nmesh_{k}"])).abs().max().item() < 1e-5, k
    assert torch.equal(tir.cpu(), torch.from_numpy(G["tir_mask"]))
    for k in range(n):
        ref, got = torch.from_numpy(G[f"path_{k}"]), pathes[k].cpu()
        assert got.shape == ref.shape, k
        err = (got - ref).norm(dim=-1)
        if k != 1:
            assert (err / ref.norm(dim=-1).clamp_min(1.0)).max().item() < 1e-4, (k, err.max().item())
        else:
            e = err.flatten()             # SDF-guided up-sampling: CDF inversion on the inner field (quantile gate)
            assert torch.quantile(e, 0.95).item() < 2e-2 and e.max().item() < 0.5, (torch.quantile(e, 0.95).item(), e.max().item())
            assert (got[:, [0, -1]] - ref[:, [0, -1]]).abs().max().item() < 1e-5


def test_ray_trace_matches_reference_sphere(net, golden):
    _check_trace(net, golden)


def test_ray_trace_matches_reference_torus():
    """Both curvature signs, rays that re-enter the object (third bounce with hits)."""
    G = np.load(os.path.join(GOLDEN, "stage2nz_torus_R96.npz"))
    net_ = make_stage2("split", mesh=torus(), thick=True).cuda()
    net_._prepare()
    # the interpolated curvature the shell geometry consumes: product definition vs the oracle-side one the reference was fed
    o, d = torch.from_numpy(G["in_o_0"]).to(DEV), torch.from_numpy(G["in_d_0"]).to(DEV)
    info, hit = net_.scene.Dintersect(o.contiguous(), d.contiguous())
    gk = info["g_k"][hit].cpu().flatten()
    ref = torch.from_numpy(G["in_gk_0"]).flatten()
    assert (gk - ref).abs().max().item() < 1e-4 * max(1.0, ref.abs().max().item())
    assert (ref > 0).any() and (ref < 0).any()
    _check_trace(net_, G)


@pytest.mark.parametrize("sph", [False, True])
def test_render_core_matches_reference(net, golden, sph):
    """sph: shader_config.sphere_direction true in both stages (shade_encode_*_var_kernel<., ., true>, 144-wide outer light)."""
    G = np.load(os.path.join(GOLDEN, "stage2nz_sph_sphere_R64.npz")) if sph else golden
    net = make_stage2("split", thick=True, sphere_direction=True).cuda() if sph else net
    o, d = torch.from_numpy(G["o"]).to(DEV), torch.from_numpy(G["d"]).to(DEV)
    pathes, converges, directions, iors, bkgr, nmesh = _lists(G)
    for mode, is_train in (("train", True), ("eval", False)):
        with torch.no_grad():
            out = net.render_core(o, d, pathes, converges, directions, bkgr, nmesh, iors, None, cos_anneal_ratio=0.2,
                                  step=10000, is_train=is_train, is_nerf=True)
        err = (out["ray_rgb"].cpu() - torch.from_numpy(G[f"{mode}_ray_rgb"])).abs().max().item()
        print(f"[NZ render_core, {mode}, sphere_direction={sph}] max |d rgb| = {err:.2e}")
        assert err < 1e-4, (mode, err)
        assert abs(out["std"].item() - float(G[f"{mode}_std"])) < 1e-6
        ge = torch.from_numpy(G[f"{mode}_gradient_error"])
        assert out["gradient_error"].shape == ge.shape and (out["gradient_error"].cpu() - ge).abs().max().item() < 1e-4
        assert out["loss_occ"].shape == (1,) and out["loss_occ"].item() == 0.0
        if not is_train:
            for key in ("normal", "specular_color", "specular_light", "specular_ref"):
                e = (out[key].cpu() - torch.from_numpy(G[f"eval_{key}"])).abs().max().item()
                assert e < 2e-4, (key, e)


def test_render_end_to_end_and_bf16(net, golden):
    G = golden
    o, d = torch.from_numpy(G["o"]).to(DEV), torch.from_numpy(G["d"]).to(DEV)
    with torch.no_grad():
        out = net.render(o, d, None, None, None, None, -1, 0.2, is_train=True, step=10000, is_nerf=True)
    ref = torch.from_numpy(G["train_ray_rgb"])
    err = (out["ray_rgb"].cpu() - ref).abs().max().item()
    print(f"[NZ render end to end, split] max |d rgb| = {err:.2e}")
    assert err < 2e-3, err
    assert torch.equal(out["tir_mask"].cpu(), torch.from_numpy(G["tir_mask"]))
    net16 = make_stage2("bf16", thick=True).cuda()
    with torch.no_grad():
        out16 = net16.render(o, d, None, None, None, None, -1, 0.2, is_train=True, step=10000, is_nerf=True)
    e16 = (out16["ray_rgb"].cpu() - ref).abs().max().item()
    print(f"[NZ render end to end, bf16] max |d rgb| = {e16:.2e}")
    assert e16 < 2e-2, e16


@pytest.mark.parametrize("case", ["sphere", "sphere_sph", "torus"])
def test_parameter_gradients_match_reference(golden, case):
    """Trainer loss backward through the product's own trace (hit kernels, IoR / thickness MLPs, shell geometry) and
    render_core on the sample parameters of the reference's trace, against the reference's autograd: every parameter
    tensor, strided samples relative to the tensor's largest entry and the norm.  Gates: 1e-2 on every tensor (single-entry
    outliers stated below), 2e-3 on at least 90 % of them (the stage-1 material layers see ~55 surface hits behind ReLU / clamp kinks, cf. the zero-thickness
    test), and the IoR / thickness networks -- whose gradient exists only through the path geometry -- within 5e-3."""
    sph = case == "sphere_sph"
    if case == "torus":
        # vertex curvatures differ across a triangle: the shell's curvature radius moves with the refracted path
        # (Stage2Renderer._curvature_with_graph; DiffRender.py:113-116), third-bounce hits, both curvature signs
        G = np.load(os.path.join(GOLDEN, "stage2nz_torus_R96.npz"))
        GG = np.load(os.path.join(GOLDEN, "stage2nz_torus_grads_R96.npz"))
        net_ = make_stage2("split", mesh=torus(), thick=True).cuda()
    else:
        G = np.load(os.path.join(GOLDEN, "stage2nz_sph_sphere_R64.npz")) if sph else golden
        GG = np.load(os.path.join(GOLDEN, "stage2nz_sph_grads_R64.npz" if sph else "stage2nz_grads_R64.npz"))
        net_ = make_stage2("split", thick=True, sphere_direction=sph).cuda()
    o, d = torch.from_numpy(G["o"]).to(DEV), torch.from_numpy(G["d"]).to(DEV)
    # sample parameters of segment 1 as the reference drew them: z = |p - start| / |end - start|
    p1 = torch.from_numpy(G["path_1"]).to(DEV)
    z1 = ((p1 - p1[:, :1]).norm(dim=-1) / (p1[:, -1:] - p1[:, :1]).norm(dim=-1)).contiguous()
    net_.zero_grad()
    prepared = net_._prepare()
    pathes, converges, directions, iors, bkgr, nmesh, tir = net_.ray_trace(o, d, None, prepared=prepared, trace={"z_1": z1})
    assert pathes[1].requires_grad and directions[1].requires_grad
    assert (pathes[1].detach() - p1).abs().max().item() < 2e-5
    out = net_.render_core(o, d, pathes, converges, directions, bkgr, nmesh, iors, None, cos_anneal_ratio=0.2, step=10000,
                           is_train=True, is_nerf=True, prepared=prepared)
    gt, tm = torch.from_numpy(GG["gt"]).to(DEV), tir
    loss = net_.compute_rgb_loss(out["ray_rgb"] * tm, gt * tm).mean() + (0.02 * out["gradient_error"]).mean()
    loss.backward()
    assert abs(loss.item() - float(GG["loss"])) < 1e-4, (loss.item(), float(GG["loss"]))
    named = dict(net_.named_parameters())
    rep = []
    for key in GG.files:
        if not key.startswith("grad/"):
            continue
        name = key[5:]
        ref, ref_norm = torch.from_numpy(GG[key]), float(GG["gradnorm/" + name])
        p = named[name]
        if ref_norm == 0.0:
            assert p.grad is None or p.grad.abs().max().item() < 1e-9, name
            continue
        assert p.grad is not None, f"no gradient for {name}"
        g = p.grad.detach().reshape(-1).cpu()
        idx = torch.linspace(0, g.numel() - 1, min(g.numel(), 64)).long()
        scale = max(ref.abs().max().item(), ref_norm / max(g.numel(), 1) ** 0.5)
        rep.append((name, (g[idx] - ref).abs().max().item() / scale, abs(p.grad.double().norm().item() - ref_norm) / ref_norm))
    rep.sort(key=lambda r: -r[1])
    geo = [r for r in rep if r[0].startswith(("IORs_pred", "thickness_pred"))]
    print(f"[NZ gradients, split, {case}] {len(rep)} tensors; worst (name, sampled rel. error, norm rel. error):")
    for r in rep[:6]:
        print("   %-60s %.2e %.2e" % r)
    print("   IoR / thickness networks:", [(n_, round(a, 6), round(b, 6)) for n_, a, b in geo])
    assert len(rep) >= 250 and len(geo) == 24, (len(rep), len(geo))
    assert all(a < 5e-3 and b < 1e-2 for _, a, b in geo), geo
    # single-entry outliers: a tensor whose NORM agrees to 1e-3 may have one strided sample off by up to 3e-2 of the largest
    # entry (measured: torus case, outer_light.0.weight_v, one of 64 samples at 2.1e-2 with the norm within 2.7e-4 -- a
    # first-layer column fed by a near-zero IDE input); at most 1 % of the tensors
    loose = [r for r in rep if r[1] >= 1e-2]
    assert all(a < 3e-2 and b < 1e-3 for _, a, b in loose) and len(loose) <= 0.01 * len(rep), loose
    assert all(b < 2e-2 for _, a, b in rep), rep[:4]
    assert sum(a < 2e-3 for _, a, _ in rep) >= 0.9 * len(rep)


# ------------------------------------------------------------------------------------------------ stage 1 of the NZ module
def _nz_stage1(precision, sph):
    from nu_nerf_b200.renderer import name2renderer
    from nu_nerf_b200.renderer_zerothick import load_default_cfg
    torch.manual_seed(0)
    cfg = load_default_cfg()
    cfg["precision"] = precision
    cfg["shader_config"] = {"sphere_direction": bool(sph), "human_light": False}
    return name2renderer["shape"](cfg, training=False).cuda()


@pytest.mark.parametrize("sph", [False, True])
def test_stage1_render_core_matches_reference(sph):
    """NeROShapeRenderer of network/renderer.py (NZ:738-859) against the unmodified reference (tests/golden/stage1nz_*.npz):
    outputs 1e-4 (loss_normal, the candidate subset of color_bkgr / color_spec -- with sphere_direction the probe sees
    [IDE(d) | IDE(exit direction)] -- included), and every parameter gradient of a loss that contains the normal-orientation,
    mask and outer-regularisation terms: strided samples within 1e-2 of the tensor's largest entry on every tensor, 2e-3 on
    at least 90 % of them, norms within 2e-2 (64 rays: a tensor's gradient rests on few samples behind ReLU / clamp kinks,
    cf. test_engine_gpu._check_golden_gradients)."""
    G = np.load(os.path.join(GOLDEN, "stage1nz_sph_R64.npz" if sph else "stage1nz_R64.npz"))
    T = lambda k: torch.from_numpy(G[k]).to(DEV)
    net = _nz_stage1("split", sph)
    net.zero_grad()
    out = net.render_core(T("o"), T("d"), T("z_vals"), None, cos_anneal_ratio=float(G["cos_anneal"]), step=int(G["step"]),
                          is_train=True, is_nerf=True)
    errs = {}
    for k in ("ray_rgb", "acc", "loss_normal", "color_bkgr", "color_spec", "transmission", "metallic"):
        ref = torch.from_numpy(G["out_" + k])
        assert out[k].shape == ref.shape, (k, out[k].shape, ref.shape)
        errs[k] = (out[k].detach().cpu() - ref).abs().max().item()
    print(f"[NZ stage 1, sphere_direction={sph}] max abs errors: " + ", ".join(f"{k} {v:.2e}" for k, v in errs.items()))
    assert all(v < 1e-4 for v in errs.values()), errs
    assert 0 < out["color_spec"].shape[0] < 64                       # the candidate subset is a proper subset here
    assert (out["gradient_error"].detach().cpu() - torch.from_numpy(G["out_gradient_error"])).abs().max().item() < 2e-3
    loss = net.compute_rgb_loss(out["ray_rgb"], T("gt")).mean() + (0.1 * out["gradient_error"]).mean() \
        + out["loss_normal"].mean() \
        + 0.5 * torch.nn.functional.mse_loss(out["color_bkgr"].flatten(), out["color_spec"].flatten()) \
        + 0.5 * torch.nn.functional.l1_loss(T("masks"), out["acc"], reduction="mean")
    loss.backward()
    assert abs(loss.item() - float(G["loss"])) < 1e-4, (loss.item(), float(G["loss"]))
    rep = []
    for name, p in net.named_parameters():
        if "grad/" + name not in G.files:
            continue
        ref, ref_norm = torch.from_numpy(G["grad/" + name]), float(G["gradnorm/" + name])
        if ref_norm == 0.0:
            continue
        assert p.grad is not None, name
        g = p.grad.detach().reshape(-1).cpu()
        idx = torch.linspace(0, g.numel() - 1, min(g.numel(), 64)).long()
        scale = max(ref.abs().max().item(), ref_norm / max(g.numel(), 1) ** 0.5)
        rep.append((name, (g[idx] - ref).abs().max().item() / scale, abs(p.grad.double().norm().item() - ref_norm) / ref_norm))
    rep.sort(key=lambda r: -r[1])
    print(f"[NZ stage 1 gradients, sphere_direction={sph}] {len(rep)} tensors; worst (name, sampled rel. error, norm rel. error):")
    for r in rep[:6]:
        print("   %-60s %.2e %.2e" % r)
    assert len(rep) > 100
    assert all(a < 1e-2 and b < 2e-2 for _, a, b in rep), rep[:4]
    assert sum(a < 2e-3 for _, a, _ in rep) >= 0.9 * len(rep)


def test_stage1_train_step_adds_loss_mask():
    net = _nz_stage1("bf16", False)
    net.is_nerf = True
    G = np.load(os.path.join(GOLDEN, "stage1nz_R64.npz"))
    T = lambda k: torch.from_numpy(G[k]).to(DEV)
    net.set_ray_source(lambda step, n: {"rays_o": T("o"), "rays_d": T("d"), "rgbs": T("gt"), "masks": T("masks")})
    out = net.train_step(10000)
    ref = torch.nn.functional.l1_loss(T("masks"), out["acc"])
    assert out["loss_mask"].shape == () and abs(out["loss_mask"].item() - ref.item()) < 1e-7
    assert out["loss_normal"].shape == (64, 1) and "loss_rgb" in out
    (out["loss_mask"] + out["loss_normal"].mean()).backward()
    assert net.sdf_network.lin0.weight_v.grad is not None and net.sdf_network.lin0.weight_v.grad.abs().sum().item() > 0


def test_shell_kernels_match_the_torch_restatement():
    """csrc/shell.cu (one launch per bounce, hand-derived adjoint) against nu_nerf_b200/shell.shell_bounce and its autograd
    on the reference's recorded bounce inputs (torus: entering, leaving, re-entering; both curvature signs)."""
    from nu_nerf_b200.renderer import shell_bounce_kernels
    from nu_nerf_b200.shell import shell_bounce, signed_normal
    G = np.load(os.path.join(GOLDEN, "stage2nz_torus_R96.npz"))
    T = lambda k: torch.from_numpy(G[k]).to(DEV)
    for k in range(int(G["n_bounces"])):
        idx = T(f"in_hit_{k}").bool().flatten().nonzero().squeeze(1)
        if idx.numel() == 0:
            continue
        inside = k % 2 == 1
        base = [T(f"in_x_{k}"), signed_normal(T(f"in_n_{k}"), inside), T(f"in_d_{k}")[idx], T(f"in_gk_{k}").reshape(-1, 1),
                T(f"in_ior_{k}").reshape(-1, 1), T(f"in_thick_{k}").reshape(-1, 1)]
        res = []
        for fn in (shell_bounce_kernels, shell_bounce):
            leaves = [t.clone().float().requires_grad_(True) for t in base]
            b = fn(*leaves, inside)
            gen = torch.Generator().manual_seed(k)
            cot = {key: torch.randn(b[key].shape, generator=gen).to(DEV) for key in ("start", "dir", "ratio", "x_mod", "normal")}
            sum((b[key] * cot[key]).sum() for key in cot).backward()
            res.append((b, [t.grad for t in leaves]))
        (bk, gk_), (bt, gt_) = res
        assert torch.equal(bk["ok"], bt["ok"]) and torch.equal(bk["tir"], bt["tir"]) and torch.equal(bk["ok_idx"], bt["ok_idx"])
        for key in ("start", "dir", "ratio", "x_mod", "normal"):
            assert (bk[key] - bt[key]).abs().max().item() <= 2e-6, (k, key)
        # (the curvature gradient is a sum of chord terms that cancel to ~1e-2 of their size: 5e-4 there, fp32 on both sides)
        for j, (a, b_) in enumerate(zip(gk_, gt_)):
            tol = 5e-4 if j == 3 else 2e-5
            assert (a - b_).abs().max().item() <= tol * max(b_.abs().max().item(), 1e-12), (k, j)


def test_edge_cases_all_rays_miss_single_ray_and_forward(net):
    """Ragged path lists of the non-zero-thickness trace: a batch whose rays all miss the outer mesh (one background segment
    of 64 inverse-depth samples, no bounce), a batch of one ray through the centre (three segments: 64 / 128 / 64 samples),
    and forward({'step'}) / forward({'eval'}) through attached sources with the reference's masks (NZ:1297-1299, :1364)."""
    from nu_nerf_b200 import feeder
    o_miss = torch.tensor([[3.0, 0.0, 0.0], [3.0, 0.1, 0.0], [0.0, 3.0, 0.2]], device=DEV)
    d_miss = torch.nn.functional.normalize(torch.tensor([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [1.0, 0.0, 0.0]], device=DEV), dim=-1)
    o_one = torch.tensor([[0.0, 0.0, 3.0]], device=DEV)
    d_one = torch.tensor([[0.0, 0.0, -1.0]], device=DEV)
    for name, o, d, shapes in (("all miss", o_miss, d_miss, [64]), ("single ray", o_one, d_one, [64, 128, 64])):
        with torch.no_grad():
            lists = net.ray_trace(o, d, None)
        assert [p.shape[1] for p in lists[0]] == shapes, (name, [p.shape for p in lists[0]])
        net.zero_grad()
        out = net.render(o, d, None, None, None, None, -1, 0.2, is_train=True, step=10000, is_nerf=True)
        assert out["ray_rgb"].shape == (o.shape[0], 3) and torch.isfinite(out["ray_rgb"]).all(), name
        assert out["tir_mask"].shape == (o.shape[0], 1) and out["loss_occ"].shape == (1,)
        (out["ray_rgb"].sum() + (0.02 * out["gradient_error"]).mean()).backward()
        g = net.stage1_network.outer_nerf.pts_linears[0].weight.grad
        assert g is not None and torch.isfinite(g).all(), name
        if name == "single ray":
            gi = net.thickness_pred.module0[0].weight_v.grad
            assert gi is not None and torch.isfinite(gi).all() and gi.abs().sum().item() > 0
    # train forward with a foreground mask: masked rays drop out of loss_rgb
    G = np.load(os.path.join(GOLDEN, "stage2nz_sphere_R64.npz"))
    o, d = torch.from_numpy(G["o"]).to(DEV), torch.from_numpy(G["d"]).to(DEV)
    gt = torch.rand(64, 3, device=DEV)
    masks = (torch.arange(64, device=DEV) % 2).float()
    net.set_ray_source(lambda step, n: {"rays_o": o, "rays_d": d, "rgbs": gt, "masks": masks})
    net.cfg["train_ray_num"] = 64
    out = net({"step": 10000})
    tm = out["tir_mask"].float() * masks[:, None]
    ref = net.compute_rgb_loss(out["ray_rgb"] * tm, gt * tm)
    assert out["loss_rgb"].shape == (64,) and (out["loss_rgb"] - ref).abs().max().item() < 1e-7
    assert abs(out["loss_rgb"][0].item() - 0.001 ** 0.5) < 1e-6                 # a masked ray: sqrt(0 + 0.001)
    # eval forward
    h, w = 4, 6
    imgs = torch.rand(1, 3, h, w, generator=torch.Generator().manual_seed(2)).to(DEV)
    K = torch.tensor([[12.0, 0, 3.0], [0, 12.0, 2.0], [0, 0, 1]])[None].to(DEV)
    c2w = torch.eye(3, 4)[None].clone()
    c2w[0, 2, 3] = 3.0
    net.set_eval_source(feeder.image_eval_source(imgs, K, c2w.to(DEV), is_nerf=True))
    out = net({"eval": True, "index": 0, "step": 10000})
    assert out["ray_rgb"].shape == (h, w, 3) and out["gt_rgb"].shape == (h, w, 3) and out["loss_rgb"].shape == (h * w,)
    assert torch.isfinite(out["ray_rgb"]).all()


def test_inner_field_occlusion_loss_matches_reference(golden):
    """NZ:2222-2230, :1580-1608 at step 20000: the occlusion-probe loss of the INNER field (predicted occlusion probability
    of color_network_inner against the hit probability of a 64 + 16 sample probe of sdf_network_inner along the reflected
    ray).  The probe goes through two CDF inversions: loss within 2e-3 relative (the stage-1 gate); its gradient reaches the
    inner_weight predictor only (12 tensors): norms within 1e-2, strided samples within 1e-2 of the largest entry."""
    G = golden
    net_ = make_stage2("split", thick=True).cuda()
    net_.cfg["occ_sdf_thresh"], net_.cfg["occ_loss_max_pn"] = 0.05, 1 << 20          # as in make_golden_nz.py
    net_._occ_perm = torch.arange(1)              # the reference's exact (non sync-free) selection path; no sub-sampling here
    o, d = torch.from_numpy(G["o"]).to(DEV), torch.from_numpy(G["d"]).to(DEV)
    pathes, converges, directions, iors, bkgr, nmesh = _lists(G)
    net_.zero_grad()
    out = net_.render_core(o, d, pathes, converges, directions, bkgr, nmesh, iors, None, cos_anneal_ratio=0.2, step=20000,
                           is_train=True, is_nerf=True)
    ref = float(G["occ20_loss_occ"][0])
    got = out["loss_occ"].mean().item()
    print(f"[NZ inner occlusion loss] {got:.6f} vs reference {ref:.6f}")
    assert ref > 1e-3 and abs(got - ref) < 2e-3 * ref + 1e-5, (got, ref)
    assert (out["ray_rgb"].detach().cpu() - torch.from_numpy(G["occ20_ray_rgb"])).abs().max().item() < 1e-4
    out["loss_occ"].mean().backward()
    named = dict(net_.named_parameters())
    keys = [k for k in G.files if k.startswith("occ20_grad/")]
    assert len(keys) == 12 and all("color_network_inner.inner_weight" in k for k in keys)
    for key in keys:
        name = key[len("occ20_grad/"):]
        ref_g, ref_norm = torch.from_numpy(G[key]), float(G["occ20_gradnorm/" + name])
        g = named[name].grad.detach().reshape(-1).cpu()
        idx = torch.linspace(0, g.numel() - 1, min(g.numel(), 64)).long()
        scale = max(ref_g.abs().max().item(), ref_norm / max(g.numel(), 1) ** 0.5)
        assert (g[idx] - ref_g).abs().max().item() / scale < 1e-2, name
        assert abs(named[name].grad.double().norm().item() - ref_norm) / ref_norm < 1e-2, name
    # before occ_loss_step / disabled: zeros(1)
    with torch.no_grad():
        out0 = net_.render_core(o, d, pathes, converges, directions, bkgr, nmesh, iors, None, cos_anneal_ratio=0.2,
                                step=10000, is_train=True, is_nerf=True)
    assert out0["loss_occ"].shape == (1,) and out0["loss_occ"].item() == 0.0

"""Kernel SOURCE on the CPU (`-m "not gpu"`): the per-point / per-ray math of the CUDA kernels lives in
nu_nerf_b200/csrc/pointwise.cuh as __host__ __device__ functions; tests/hostsim/shell_host.cpp builds it for the host (g++)
and this file checks it -- values and the hand-derived adjoints -- against (1) the torch restatement of the
non-zero-thickness bounce nu_nerf_b200/shell.py, which is itself pinned bounce by bounce to the UNMODIFIED reference's
ray_trace (network/renderer.py:1610-2148) on the inputs the reference saw (tests/golden/stage2nz_*.npz, made by
make_golden_nz.py), (2) the reference's own expressions restated in torch (sphere direction field.py:447-465, sdf -> alpha
renderer_zerothick.py:669-684, NeRF++ activations :515-516 / :691-692, shading directions field.py:686-689) and their autograd
in fp64, and (3) the oracle's integrated directional encoding.  The GPU suite then only has to show that the launched
kernels compute what this source says."""
import os

import numpy as np
import pytest
import torch

from nu_nerf_b200.shell import shell_bounce, signed_normal, outside_depths

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 2e-6


def _t(a):
    return torch.from_numpy(np.asarray(a))


@pytest.mark.parametrize("name", ["stage2nz_sphere_R64.npz", "stage2nz_torus_R96.npz"])
def test_shell_bounce_matches_reference_trace(name):
    g = np.load(os.path.join(G, name))
    n_b, n_seg = int(g["n_bounces"]), int(g["n_segments"])
    tirs, conv_rows = [], []
    seen_signs = set()
    for k in range(n_b):
        hit = _t(g[f"in_hit_{k}"]).bool().flatten()
        if hit.sum() == 0:
            assert k == n_b - 1
            break
        hit_idx = hit.nonzero().squeeze(1)
        d_all = _t(g[f"in_d_{k}"])
        inside = k % 2 == 1
        out = shell_bounce(_t(g[f"in_x_{k}"]), signed_normal(_t(g[f"in_n_{k}"]), inside), d_all[hit_idx], _t(g[f"in_gk_{k}"]),
                           _t(g[f"in_ior_{k}"]).reshape(-1, 1), _t(g[f"in_thick_{k}"]).reshape(-1, 1), inside)
        seen_signs |= set(np.sign(g[f"in_gk_{k}"]).flatten().tolist())
        conv = torch.zeros(hit.shape[0], dtype=torch.bool)
        conv[hit_idx] = out["ok"]
        assert torch.equal(conv, _t(g[f"converge_{k}"]).bool().flatten()), f"bounce {k}: pass mask"
        tir = torch.ones(hit.shape[0], dtype=torch.bool)
        tir[hit_idx] = out["tir"]
        tirs.append(tir)
        conv_rows.append(conv)
        # the hit point the segment ends at (pulled back onto the inner face when leaving the object)
        end = _t(g[f"path_{k}"])[hit_idx, -1, :]
        assert (out["x_mod"] - end).abs().max().item() <= 4e-6, f"bounce {k}: hit point"
        if out["ok_idx"].numel() == 0:
            break
        assert (out["normal"] - _t(g[f"nmesh_{k}"])).abs().max().item() <= TOL
        assert (out["ratio"] - _t(g[f"ior_{k}"])).abs().max().item() <= TOL
        assert (out["dir"] - _t(g[f"dir_{k + 1}"])).abs().max().item() <= TOL, f"bounce {k}: next direction"
        if k + 1 < n_b:                                    # the origin the reference traced the next segment from
            assert (out["start"] - _t(g[f"in_o_{k + 1}"])).abs().max().item() <= TOL, f"bounce {k}: next origin"
    for i in range(len(tirs) - 1, 0, -1):                 # NZ:2063-2064
        tirs[i - 1][conv_rows[i - 1]] &= tirs[i]
    assert torch.equal(tirs[0], _t(g["tir_mask"]).bool().flatten())
    if "torus" in name:
        assert {-1.0, 1.0} <= seen_signs                 # both curvature branches exercised


def test_shell_bounce_is_differentiable_and_handles_empty():
    g = np.load(os.path.join(G, "stage2nz_torus_R96.npz"))
    hit_idx = _t(g["in_hit_1"]).bool().flatten().nonzero().squeeze(1)
    ior = _t(g["in_ior_1"]).reshape(-1, 1).clone().requires_grad_(True)
    th = _t(g["in_thick_1"]).reshape(-1, 1).clone().requires_grad_(True)
    x = _t(g["in_x_1"]).clone().requires_grad_(True)
    out = shell_bounce(x, signed_normal(_t(g["in_n_1"]), True), _t(g["in_d_1"])[hit_idx], _t(g["in_gk_1"]), ior, th, True)
    (out["dir"].sum() + out["start"].square().sum()).backward()
    for t in (ior, th, x):
        assert t.grad is not None and torch.isfinite(t.grad).all() and t.grad.abs().sum() > 0
    e = torch.zeros(0, 3)
    out = shell_bounce(e, e, e, torch.zeros(0, 1), torch.zeros(0, 1), torch.zeros(0, 1), False)
    assert out["ok_idx"].numel() == 0 and out["start"].shape == (0, 3)
    z = outside_depths("cpu")
    assert z.shape == (64,) and abs(z[0].item() - (1.0 / (1.0 - 1.0 / 65.0) + 1.0 / 64)) < 1e-6 and z[-1].item() > 1000.0


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    """pw::shell_bounce_fwd / _bwd -- the source the CUDA kernels of csrc/shell.cu compile -- built for the host (g++)."""
    import ctypes
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = str(tmp_path_factory.mktemp("hostsim") / "libshell_host.so")
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-I", os.path.join(root, "nu_nerf_b200", "csrc"),
                    "-o", so, os.path.join(root, "tests", "hostsim", "shell_host.cpp")], check=True)
    return ctypes.CDLL(so)


@pytest.mark.parametrize("name", ["stage2nz_sphere_R64.npz", "stage2nz_torus_R96.npz"])
def test_kernel_source_matches_torch_restatement_and_its_autograd(host_lib, name):
    """The bounce kernels' math (forward and hand-derived adjoint) on the reference's recorded inputs, every bounce: masks
    equal, values within 1e-6, gradients of random cotangents within 2e-5 of the largest entry against torch autograd."""
    import ctypes
    fp = ctypes.POINTER(ctypes.c_float)
    ptr = lambda a: a.ctypes.data_as(fp)
    g = np.load(os.path.join(G, name))
    checked = 0
    for k in range(int(g["n_bounces"])):
        hit = _t(g[f"in_hit_{k}"]).bool().flatten()
        idx = hit.nonzero().squeeze(1)
        if idx.numel() == 0:
            continue
        inside = k % 2 == 1
        leaves = [_t(g[f"in_x_{k}"]), signed_normal(_t(g[f"in_n_{k}"]), inside), _t(g[f"in_d_{k}"])[idx],
                  _t(g[f"in_gk_{k}"]).reshape(-1, 1), _t(g[f"in_ior_{k}"]).reshape(-1, 1), _t(g[f"in_thick_{k}"]).reshape(-1, 1)]
        leaves = [t.clone().float().requires_grad_(True) for t in leaves]
        M = leaves[0].shape[0]
        inp = np.ascontiguousarray(torch.cat([t.detach() for t in leaves], 1).numpy(), dtype=np.float32)
        out = np.zeros((M, 12), np.float32)
        host_lib.shell_fwd(ptr(inp), M, int(inside), ptr(out))
        b = shell_bounce(*leaves, inside)
        oi = b["ok_idx"].numpy()
        assert np.array_equal(out[:, 0] > 0, b["ok"].numpy()) and np.array_equal(out[:, 1] > 0, b["tir"].numpy())
        err = lambda a, t: float(np.abs(a - t.detach().numpy()).max()) if a.size else 0.0
        assert err(out[:, 2:5], b["x_mod"]) <= 1e-6 and err(out[oi, 5:8], b["start"]) <= 1e-6
        assert err(out[oi, 8:11], b["dir"]) <= 1e-6 and err(out[oi, 11:12], b["ratio"]) <= 1e-6
        gen = torch.Generator().manual_seed(k)
        K = oi.shape[0]
        gs, gd, gr = torch.randn(K, 3, generator=gen), torch.randn(K, 3, generator=gen), torch.randn(K, 1, generator=gen)
        gx = torch.randn(M, 3, generator=gen)
        ((b["start"] * gs).sum() + (b["dir"] * gd).sum() + (b["ratio"] * gr).sum() + (b["x_mod"] * gx).sum()).backward()
        cot = np.ascontiguousarray(np.concatenate([gs.numpy(), gd.numpy(), gr.numpy(), gx.numpy()[oi]], 1), dtype=np.float32)
        din = np.zeros((K, 12), np.float32)
        host_lib.shell_bwd(ptr(np.ascontiguousarray(inp[oi])), K, int(inside), ptr(cot), ptr(din))
        ref = torch.cat([t.grad for t in leaves], 1).numpy()[oi]
        for lo, hi in ((0, 3), (3, 6), (6, 9), (9, 10), (10, 11), (11, 12)):
            scale = max(float(np.abs(ref[:, lo:hi]).max()), 1e-12)
            # (the curvature gradient, column 9, is a sum of chord terms that cancel to ~1e-2 of their size)
            assert float(np.abs(din[:, lo:hi] - ref[:, lo:hi]).max()) / scale <= (2e-4 if lo == 9 else 2e-5), (k, lo)
        checked += 1
    assert checked >= 2


def test_sphere_direction_kernel_source_matches_reference_functions(host_lib):
    """pw::sphere_dir_fwd / _bwd (the exit point of the ray (p, u) on the unit sphere, encoded a second time by the
    `sphere_direction` shaders) against the reference's own offset_points_to_sphere + get_sphere_intersection + F.normalize
    (field.py:447-465, :641-644), restated in torch, and against their autograd -- points inside AND outside radius 0.999."""
    import ctypes
    import torch.nn.functional as F
    fp = ctypes.POINTER(ctypes.c_float)
    ptr = lambda a: a.ctypes.data_as(fp)
    gen = torch.Generator().manual_seed(4)
    M = 400
    p = torch.randn(M, 3, generator=gen) * torch.where(torch.arange(M) % 2 == 0, 0.5, 1.2)[:, None]
    u = F.normalize(torch.randn(M, 3, generator=gen), dim=-1)
    p.requires_grad_(True)
    u.requires_grad_(True)

    def reference(points, dirs):                       # field.py:447-465 as written
        norm = torch.norm(points, dim=-1)
        mask = norm > 0.999
        pts = torch.clone(points)
        pts[mask] = pts[mask] / norm[mask].unsqueeze(-1) * 0.999
        dtx = torch.sum(pts * dirs, dim=-1, keepdim=True)
        xtx = torch.sum(pts ** 2, dim=-1, keepdim=True)
        dist = -dtx + torch.sqrt(dtx ** 2 - xtx + 1 + 1e-6)
        return F.normalize(pts + dirs * dist, dim=-1)
    q_ref = reference(p, u)
    assert (p.detach().norm(dim=-1) > 0.999).any() and (p.detach().norm(dim=-1) < 0.999).any()
    pn, un = (np.ascontiguousarray(t.detach().numpy(), dtype=np.float32) for t in (p, u))
    q = np.zeros((M, 3), np.float32)
    host_lib.sphere_dir(ptr(pn), ptr(un), M, ptr(q))
    assert np.abs(q - q_ref.detach().numpy()).max() <= 5e-6      # points just outside the sphere: sqrt of a cancelling sum
    from nu_nerf_b200.engine import sphere_exit_dir   # the [R,3] torch glue of the per-ray specular probe
    assert (sphere_exit_dir(p.detach(), u.detach()) - q_ref.detach()).abs().max().item() <= 5e-6
    cot = torch.randn(M, 3, generator=gen)
    (q_ref * cot).sum().backward()
    du, dp = np.zeros((M, 3), np.float32), np.zeros((M, 3), np.float32)
    host_lib.sphere_dir_bwd(ptr(pn), ptr(un), ptr(np.ascontiguousarray(cot.numpy())), M, ptr(du), ptr(dp))
    for got, ref in ((du, u.grad.numpy()), (dp, p.grad.numpy())):
        assert np.abs(got - ref).max() <= 2e-5 * max(np.abs(ref).max(), 1.0)


def test_shading_mix_kernel_source_with_the_specinner_refraction_clamp(host_lib):
    """pw::shade_mix_fwd / _bwd with a refraction-light clamp different from light_exp_max (AppShadingNetwork_SpecInner:
    exp(min(x, -0.2)), field.py:1373): the refraction term saturates at exp(-0.2) while the other lights keep exp_max = 5,
    and the hand-derived backward agrees with central differences of the forward (smooth region: no clamp / LUT-cell edges)."""
    import ctypes
    fp = ctypes.POINTER(ctypes.c_float)
    ptr = lambda a: a.ctypes.data_as(fp)
    lut = np.load(os.path.join(G, "fg_lut_reference.npz"))["FG_LUT"].astype(np.float32).reshape(-1)
    rng = np.random.default_rng(5)
    f = ctypes.c_float

    def fwd(v, er):
        out = np.zeros(6, np.float32)
        host_lib.mix_fwd(ptr(np.ascontiguousarray(v, dtype=np.float32)), ptr(lut), f(5.0), f(er), ptr(out))
        return out
    checked = 0
    for trial in range(40):
        v = rng.normal(0.0, 0.6, 26).astype(np.float32)
        v[25] = rng.uniform(0.15, 0.85)                         # NoV inside (0, 1)
        v[22:25] = rng.uniform(-1.5, -0.4, 3) if trial % 2 == 0 else rng.uniform(0.2, 1.5, 3)    # below / above the clamp
        # saturation: raising a clamped refraction head does not change the colour, with the shared clamp (5.0) it does
        hi = v.copy()
        hi[22:25] += 0.5
        if trial % 2 == 1:
            assert np.array_equal(fwd(v, -0.2), fwd(hi, -0.2)) and not np.array_equal(fwd(v, 5.0), fwd(hi, 5.0))
        cot = rng.normal(size=3).astype(np.float32)
        d_in = np.zeros(26, np.float32)
        host_lib.mix_bwd(ptr(v), ptr(lut), f(5.0), f(-0.2), ptr(cot), f(0.3), f(-0.2), ptr(d_in))
        if trial % 2 == 1:
            assert np.all(d_in[22:25] == 0.0)                    # clamped heads receive no gradient
        h = 2e-3
        for j in range(26):
            if j == 1 or j == 25:
                continue                                         # roughness / NoV move the bilinear LUT taps: piecewise
            a, b = v.copy(), v.copy()
            a[j] += h
            b[j] -= h
            oa, ob = fwd(a, -0.2).astype(np.float64), fwd(b, -0.2).astype(np.float64)
            fd = (np.dot(cot, oa[:3] - ob[:3]) + 0.3 * (oa[3] - ob[3]) - 0.2 * (oa[4] - ob[4])) / (2 * h)
            assert abs(fd - d_in[j]) <= 2e-3 * (1.0 + abs(fd)), (trial, j, fd, d_in[j])
            checked += 1
    assert checked > 850


def test_ide_kernel_source_matches_the_oracle_and_its_autograd(host_lib):
    """pw::ide_fwd / ide_bwd -- the integrated directional encoding every shading row is built from (shade_encode_* kernels,
    the specular probe) -- with the table the library uploads to constant memory, against oracle/nunerf_oracle.ide (pinned to
    the reference's goldens; utils/ref_utils.py:85-114), unit directions, roughness in [0, 1].  The degree-16 block is
    ill-conditioned in fp32 on BOTH sides (coefficients ~1e5 with cancellation near |z| = 1 at roughness 0: the reference's
    own fp32 evaluation is 3.6e-3 off its fp64 value), so every block is compared with the fp64 evaluation: l <= 8 within
    2e-5, l = 16 within 1e-2 and no worse than 1.5 x the fp32 oracle.  Gradients: against fp64 autograd, cotangents on the
    l <= 8 columns (2e-4 of the largest entry)."""
    import ctypes
    import sys
    import torch.nn.functional as F
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import nunerf_oracle as orc
    fp = ctypes.POINTER(ctypes.c_float)
    ptr = lambda a: a.ctypes.data_as(fp)
    gen = torch.Generator().manual_seed(6)
    M = 300
    xyz = F.normalize(torch.randn(M, 3, generator=gen), dim=-1)
    k = torch.rand(M, 1, generator=gen)
    k[::5] = 0.0
    k[1::5] = 1.0
    ref32 = orc.ide(xyz, k).numpy()
    x64, k64 = xyz.double().requires_grad_(True), k.double().requires_grad_(True)
    ref64 = orc.ide(x64, k64)
    xn = np.ascontiguousarray(xyz.numpy(), dtype=np.float32)
    kn = np.ascontiguousarray(k.numpy().reshape(-1), dtype=np.float32)
    out = np.zeros((M, 72), np.float32)
    host_lib.ide_host(ptr(xn), ptr(kn), M, ptr(out))
    r64 = ref64.detach().numpy()
    low = [c for c in range(72) if c % 36 < 19]                   # l = 1, 2, 4, 8
    high = [c for c in range(72) if c % 36 >= 19]                 # l = 16
    assert np.abs(out[:, low] - r64[:, low]).max() <= 2e-5
    e_host, e_orc = np.abs(out[:, high] - r64[:, high]).max(), np.abs(ref32[:, high] - r64[:, high]).max()
    assert e_host <= 1e-2 and e_host <= 1.5 * e_orc + 1e-5, (e_host, e_orc)
    cot = torch.randn(M, 72, generator=gen)
    cot[:, high] = 0.0
    (ref64 * cot.double()).sum().backward()
    dx, dk = np.zeros((M, 3), np.float32), np.zeros(M, np.float32)
    host_lib.ide_bwd_host(ptr(xn), ptr(kn), ptr(np.ascontiguousarray(cot.numpy())), M, ptr(dx), ptr(dk))
    for got, want in ((dx, x64.grad.numpy()), (dk, k64.grad.numpy().reshape(-1))):
        assert np.abs(got - want).max() <= 2e-4 * max(1.0, float(np.abs(want).max())), np.abs(got - want).max()


@pytest.mark.parametrize("inv_s,anneal", [(20.0, 0.2), (300.0, 1.0), (64.0, 0.0)])
def test_sdf_alpha_kernel_source_matches_the_reference_expression(host_lib, inv_s, anneal):
    """pw::sdf_alpha_fwd / _bwd (sdf_alpha_fwd/bwd_kernel) against compute_sdf_alpha as the reference writes it
    (renderer_zerothick.py:669-684) + the eikonal term (:769) and their torch autograd, at the init-like and a trained-like
    sharpness: alpha / eikonal within 2e-6, every gradient (sdf, SDF gradient, interval length, ray direction, inv_s) within
    1e-4 of the largest entry, fp64 on the torch side."""
    import ctypes
    import torch.nn.functional as F
    fp = ctypes.POINTER(ctypes.c_float)
    ptr = lambda a: a.ctypes.data_as(fp)
    f = ctypes.c_float
    gen = torch.Generator().manual_seed(int(inv_s))
    M = 2000
    sdf = (torch.randn(M, generator=gen) * (3.0 / inv_s)).double().requires_grad_(True)
    g = (F.normalize(torch.randn(M, 3, generator=gen), dim=-1) * (0.7 + 0.6 * torch.rand(M, 1, generator=gen))).double().requires_grad_(True)
    dist = (0.002 + 0.03 * torch.rand(M, generator=gen)).double().requires_grad_(True)
    dirs = F.normalize(torch.randn(M, 3, generator=gen), dim=-1).double().requires_grad_(True)
    s = torch.tensor(float(inv_s), dtype=torch.float64, requires_grad=True)
    true_cos = (dirs * g).sum(-1)
    iter_cos = -(F.relu(-true_cos * 0.5 + 0.5) * (1.0 - anneal) + F.relu(-true_cos) * anneal)
    en, ep = sdf + iter_cos * dist * 0.5, sdf - iter_cos * dist * 0.5
    pc, nc = torch.sigmoid(ep * s), torch.sigmoid(en * s)
    alpha = ((pc - nc + 1e-5) / (pc + 1e-5)).clip(0.0, 1.0)
    gerr = (torch.linalg.norm(g, ord=2, dim=-1) - 1.0) ** 2
    inp = torch.cat([sdf[:, None], g, dist[:, None], dirs, torch.zeros(M, 1)], 1).detach().float().contiguous().numpy()
    out = np.zeros((M, 2), np.float32)
    host_lib.sdf_alpha_host(ptr(inp), M, f(inv_s), f(anneal), ptr(out))
    assert np.abs(out[:, 0] - alpha.detach().numpy()).max() <= 2e-6 * max(1.0, inv_s / 50.0)
    assert np.abs(out[:, 1] - gerr.detach().numpy()).max() <= 2e-6
    cot = torch.randn(M, 2, generator=gen)
    ((alpha * cot[:, 0].double()).sum() + (gerr * cot[:, 1].double()).sum()).backward()
    d = np.zeros((M, 9), np.float32)
    host_lib.sdf_alpha_bwd_host(ptr(inp), M, f(inv_s), f(anneal), ptr(np.ascontiguousarray(cot.numpy())), ptr(d))
    interior = ((alpha.detach() > 1e-6) & (alpha.detach() < 1 - 1e-6)).numpy()          # away from the clip edges
    for got, want in ((d[:, 0], sdf.grad.numpy()), (d[:, 1:4], g.grad.numpy()), (d[:, 4], dist.grad.numpy()),
                      (d[:, 5:8], dirs.grad.numpy())):
        scale = max(1.0, float(np.abs(want).max()))
        assert np.abs(got[interior] - want[interior]).max() <= 1e-4 * scale, (np.abs(got[interior] - want[interior]).max(), scale)
    want_s = s.grad.item()
    got_s = float(d[interior, 8].astype(np.float64).sum()) + float(d[~interior, 8].astype(np.float64).sum())
    assert abs(got_s - want_s) <= 2e-4 * max(1.0, abs(want_s)), (got_s, want_s)


def test_nerf_output_and_shading_direction_kernel_sources(host_lib):
    """pw::nerf_out_fwd / _bwd against compute_density_alpha as the reference writes it (renderer_zerothick.py:515-516,
    :691-692, utils/raw_utils.py:5-12) and pw::shade_dirs / shade_dirs_bwd against the direction block of
    AppShadingNetwork.forward (field.py:686-689), values and autograd (fp64 on the torch side)."""
    import ctypes
    import torch.nn.functional as F
    fp = ctypes.POINTER(ctypes.c_float)
    ptr = lambda a: a.ctypes.data_as(fp)
    gen = torch.Generator().manual_seed(8)
    M = 1500
    # ---- NeRF++ output activation
    sigma = (torch.randn(M, generator=gen) * 4.0).double().requires_grad_(True)
    rgb = (torch.randn(M, 3, generator=gen) * 2.0 - 1.0).double().requires_grad_(True)
    dist = (0.01 + torch.rand(M, generator=gen)).double().requires_grad_(True)
    alpha = 1.0 - torch.exp(-F.softplus(sigma) * dist)
    lin = torch.exp(torch.clamp(rgb, max=5.0))
    eps = torch.finfo(torch.float32).eps
    col = torch.where(lin <= 0.0031308, 323 / 25 * lin, (211 * torch.clamp(lin, min=eps) ** (5 / 12) - 11) / 200)
    inp = torch.cat([sigma[:, None], rgb, dist[:, None]], 1).detach().float().contiguous().numpy()
    out = np.zeros((M, 4), np.float32)
    host_lib.nerf_out_host(ptr(inp), M, ptr(out))
    assert np.abs(out[:, 0] - alpha.detach().numpy()).max() <= 2e-6
    assert np.abs(out[:, 1:] - col.detach().numpy()).max() <= 2e-5 * float(col.detach().max())
    cot = torch.randn(M, 4, generator=gen)
    ((alpha * cot[:, 0].double()).sum() + (col * cot[:, 1:].double()).sum()).backward()
    d = np.zeros((M, 5), np.float32)
    host_lib.nerf_out_bwd_host(ptr(inp), M, ptr(np.ascontiguousarray(cot.numpy())), ptr(d))
    for got, want in ((d[:, 0], sigma.grad.numpy()), (d[:, 1:4], rgb.grad.numpy()), (d[:, 4], dist.grad.numpy())):
        assert np.abs(got - want).max() <= 1e-4 * max(1.0, float(np.abs(want).max()))
    # ---- shading directions
    g = (torch.randn(M, 3, generator=gen) * 1.3).double().requires_grad_(True)          # un-normalised SDF gradient
    rd = (F.normalize(torch.randn(M, 3, generator=gen), dim=-1) * 1.0).double().requires_grad_(True)
    n, v = F.normalize(g, dim=-1), F.normalize(-rd, dim=-1)
    r = torch.sum(v * n, -1, keepdim=True) * n * 2 - v
    nov = torch.sum(n * v, -1, keepdim=True)
    out = np.zeros((M, 10), np.float32)
    gn, rn = (np.ascontiguousarray(t.detach().float().numpy()) for t in (g, rd))
    host_lib.shade_dirs_host(ptr(gn), ptr(rn), M, ptr(out))
    ref = torch.cat([n, v, r, nov], 1).detach().numpy()
    assert np.abs(out - ref).max() <= 2e-6
    cot = torch.randn(M, 10, generator=gen)
    (torch.cat([n, v, r, nov], 1) * cot.double()).sum().backward()
    dg, drd = np.zeros((M, 3), np.float32), np.zeros((M, 3), np.float32)
    host_lib.shade_dirs_bwd_host(ptr(gn), ptr(rn), ptr(np.ascontiguousarray(cot.numpy())), M, ptr(dg), ptr(drd))
    for got, want in ((dg, g.grad.numpy()), (drd, rd.grad.numpy())):
        assert np.abs(got - want).max() <= 1e-4 * max(1.0, float(np.abs(want).max())), np.abs(got - want).max()

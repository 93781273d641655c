"""Non-zero-thickness bounce geometry (nu_nerf_b200/shell.shell_bounce) against the UNMODIFIED reference's
ray_trace (network/renderer.py:1610-2148), bounce by bounce, on the inputs the reference itself saw (hit point, interpolated
normal and Gaussian curvature, IoR / thickness network outputs): tests/golden/stage2nz_*.npz, made by make_golden_nz.py.
fp32 on both sides with the reference's operation order: the gate is 2e-6 (a few ulp of O(1) quantities)."""
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

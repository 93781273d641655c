"""Host logic of the non-zero-thickness Stage2Renderer (nu_nerf_b200/renderer.py) on CPU: bit-identical initialisation against
the unmodified network/renderer.py, and the bookkeeping of ray_trace (per-bounce lists, pass / TIR masks, segment sample
layout) against the reference's own trace with the device kernels replaced by the values the reference recorded
(tests/golden/stage2nz_*.npz: hit records of Scene.Dintersect, IoR / thickness network outputs).  The kernels themselves
are checked on the GPU (tests/test_renderer_nz_gpu.py)."""
import os
import types

import numpy as np
import pytest
import torch

from conftest import GOLDEN, make_stage2


def _fp(t):
    from test_oracle_golden import _fp as f
    return f(t)


@pytest.mark.parametrize("sph", [False, True])
def test_nz_initialisation_is_bit_identical_to_the_reference(sph):
    net = make_stage2(thick=True, sphere_direction=sph)
    sd = net.state_dict()
    G = np.load(os.path.join(GOLDEN, "stage2nz_sph_init.npz" if sph else "stage2nz_init.npz"))
    assert len(G.files) == 562
    assert net.stage1_network.color_network.outer_light[0].weight_v.shape[1] == (144 if sph else 72)
    for k in G.files:
        assert k in sd, f"missing parameter {k}"
        assert np.array_equal(_fp(sd[k]), G[k]), k
    extra = sorted(k for k in set(sd) - set(G.files) if not k.endswith("FG_LUT"))
    assert extra == [], extra
    from nu_nerf_b200.renderer import name2renderer
    assert set(name2renderer) == {"shape", "stage2"} and isinstance(net, name2renderer["stage2"])
    assert net.color_network_inner.refrac_light.exp_max == -0.2 and net.color_network_inner.cfg["light_pos_freq"] == 8


class _FakeEngine:
    """Stands in for the device kernels with the values the reference recorded (call order = bounce order)."""

    def __init__(self, G):
        self.G, self.calls = G, 0

    def ior_forward(self, w, x):
        k, which = divmod(self.calls, 2)
        self.calls += 1
        return torch.from_numpy(self.G[f"in_ior_{k}" if which == 0 else f"in_thick_{k}"]).flatten()

    @staticmethod
    def segment_points(start, delta, z):
        return start[:, None, :] + delta[:, None, :] * z[:, :, None]

    @staticmethod
    def sdf_infer(w, pts, planes):
        return pts.norm(dim=-1) - 0.3

    @staticmethod
    def upsample_rounds(w, o, d, z, sdf, n_new, rounds):
        extra = torch.linspace(0.25, 0.75, n_new * rounds).unsqueeze(0).expand(z.shape[0], -1)
        return torch.cat([z, extra], -1).sort(-1).values


class _FakeScene:
    def __init__(self, G):
        self.G, self.k = G, 0

    def Dintersect(self, o, d):
        G, k = self.G, self.k
        self.k += 1
        assert np.allclose(o.numpy(), G[f"in_o_{k}"], atol=3e-6) and np.allclose(d.numpy(), G[f"in_d_{k}"], atol=3e-6)
        hit = torch.from_numpy(G[f"in_hit_{k}"]).bool().flatten()
        idx = hit.nonzero().squeeze(1)
        N = hit.shape[0]

        def full(a, w):
            t = torch.zeros(N, w) if w else torch.zeros(N, dtype=torch.int32)
            t[idx] = torch.from_numpy(a).reshape(-1, w).float() if w else torch.from_numpy(a).int().flatten()
            return t
        info = {"x": full(G[f"in_x_{k}"], 3), "n": full(G[f"in_n_{k}"], 3), "g_k": full(G[f"in_gk_{k}"], 1),
                "faces_ind": full(G[f"in_tri_{k}"], 0)}
        return info, hit


@pytest.mark.parametrize("name", ["stage2nz_sphere_R64.npz", "stage2nz_torus_R96.npz"])
def test_nz_ray_trace_bookkeeping_matches_reference(name, monkeypatch):
    import nu_nerf_b200.renderer as R
    G = np.load(os.path.join(GOLDEN, name))
    net = make_stage2(thick=True)
    fake = _FakeEngine(G)
    monkeypatch.setattr(R, "_engine", lambda: fake)
    net.scene = _FakeScene(G)
    net._w_thick = None
    prepared = (None, types.SimpleNamespace(sdf=None, planes=2), None)
    o, d = torch.from_numpy(G["o"]), torch.from_numpy(G["d"])
    with torch.no_grad():
        pathes, converges, directions, iors, bkgr, nmesh, tir = net.ray_trace(o, d, None, prepared=prepared)
    n = int(G["n_segments"])
    assert len(pathes) == n and len(converges) == n and len(bkgr) == n and len(directions) == n + 1
    assert len(iors) == len(nmesh) == sum(f"ior_{k}" in G.files for k in range(n))
    for k in range(n):
        assert torch.equal(converges[k], torch.from_numpy(G[f"converge_{k}"])), k
        assert torch.equal(bkgr[k], torch.from_numpy(G[f"bkgr_{k}"])), k
        ref = torch.from_numpy(G[f"path_{k}"])
        assert pathes[k].shape == ref.shape, k
        if k != 1:
            assert (pathes[k] - ref).abs().max().item() <= 1e-5 * max(1.0, ref.abs().max().item()), k
        else:                       # up-sampled segment: the first and last samples are the segment's end points
            assert (pathes[k][:, [0, -1]] - ref[:, [0, -1]]).abs().max().item() <= 4e-6
    for k in range(n + 1):
        ref = torch.from_numpy(G[f"dir_{k}"])
        assert directions[k].shape == ref.shape and (ref.numel() == 0 or (directions[k] - ref).abs().max().item() <= 2e-6), k
    for k in range(len(iors)):
        assert (iors[k] - torch.from_numpy(G[f"ior_{k}"])).abs().max().item() <= 2e-6
        assert (nmesh[k] - torch.from_numpy(G[f"nmesh_{k}"])).abs().max().item() <= 2e-6
    assert torch.equal(tir, torch.from_numpy(G["tir_mask"]))


def test_nz_rays_without_exit_are_taken_back(monkeypatch):
    """NZ:1662-1672: a ray that entered the object and misses the mesh from the inside loses its first hit."""
    import nu_nerf_b200.renderer as R
    G = dict(np.load(os.path.join(GOLDEN, "stage2nz_sphere_R64.npz")))
    lost = 3                                                        # drop the exit hit of the 4th continuing ray
    hit1 = G["in_hit_1"].copy().reshape(-1)
    hit1[lost] = 0
    G["in_hit_1"] = hit1.reshape(G["in_hit_1"].shape)
    keep = np.ones(hit1.shape[0], bool)
    keep[lost] = False
    for key in ("in_x_1", "in_n_1", "in_gk_1", "in_tri_1", "in_ior_1", "in_thick_1"):
        G[key] = G[key][keep]
    G["in_o_2"], G["in_d_2"] = G["in_o_2"][keep], G["in_d_2"][keep]
    G["in_hit_2"] = G["in_hit_2"].reshape(-1)[keep]
    net = make_stage2(thick=True)
    monkeypatch.setattr(R, "_engine", lambda: _FakeEngine(G))
    net.scene = _FakeScene(G)
    net._w_thick = None
    prepared = (None, types.SimpleNamespace(sdf=None, planes=2), None)
    with torch.no_grad():
        pathes, converges, directions, iors, bkgr, nmesh, tir = net.ray_trace(torch.from_numpy(G["o"]), torch.from_numpy(G["d"]),
                                                                             None, prepared=prepared)
    ref0 = torch.from_numpy(np.load(os.path.join(GOLDEN, "stage2nz_sphere_R64.npz"))["converge_0"]).flatten()
    row = ref0.nonzero().squeeze(1)[lost]
    assert converges[0].sum().item() == ref0.sum().item() - 1 and not converges[0][row].item()
    n1 = int(ref0.sum()) - 1
    assert pathes[1].shape[0] == n1 and directions[1].shape[0] == n1 and iors[0].shape[0] == n1 and nmesh[0].shape[0] == n1
    assert not bkgr[0][row].item()                                  # segment 0 still ends at the hit point

"""CPU check of Stage2Renderer._replay_geometry (the differentiable re-statement of ray_trace, ZT:1571-1720, used for
the gradient of IORs_pred): with the discrete decisions of the REFERENCE's trace (tests/golden/stage2_R64.npz) it must
reproduce the reference's refracted directions, mesh normals and every sampled path point, and its graph must reach the
IoR network's parameters."""
import os
from types import SimpleNamespace

import numpy as np
import torch

from conftest import GOLDEN, make_stage2, stage2_rec_from_golden, uv_sphere


def test_replay_reproduces_the_reference_trace():
    from nu_nerf_b200.tracer import angle_weighted_vertex_normals
    G = np.load(os.path.join(GOLDEN, "stage2_R64.npz"))
    net = make_stage2("split")
    V, Fc = uv_sphere(float(G["mesh_radius"]), int(G["mesh_nu"]), int(G["mesh_nv"]))
    V, Fc = torch.as_tensor(V, dtype=torch.float64), torch.as_tensor(Fc, dtype=torch.long)
    net.scene = SimpleNamespace(vertices=V, faces=Fc, normals=angle_weighted_vertex_normals(V, Fc))
    rec = stage2_rec_from_golden(G, "cpu")
    n = int(G["n_segments"])
    T = lambda k: torch.from_numpy(G[k])
    pathes, dirs = [T(f"path_{k}") for k in range(n)], [T(f"dir_{k}") for k in range(n + 1)]
    nmesh = [T(f"nmesh_{k}") for k in range(n) if f"nmesh_{k}" in G.files]
    new_p, new_d, new_n = net._replay_geometry(T("o"), T("d"), rec, pathes, dirs, nmesh, straight_through=False)
    for k in range(1, n):
        assert (new_d[k] - dirs[k]).abs().max().item() < 1e-5, k
        assert new_p[k] is not None
        err = (new_p[k] - pathes[k]).norm(dim=-1)
        assert err.max().item() < 2e-4, (k, err.max().item())      # 64 * eps at the far samples
    for k in range(len(nmesh)):
        assert (new_n[k] - nmesh[k]).abs().max().item() < 1e-5, k
    # the graph reaches every IoR parameter
    net.zero_grad()
    (new_p[1].sum() + new_p[2].square().sum()).backward()
    for name, p in net.IORs_pred.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all() and p.grad.abs().max().item() > 0, name

"""Discrete Gaussian curvature (nu_nerf_b200/tracer.discrete_gaussian_curvature, the stated replacement for the PyMesh
attribute DiffRender.Scene reads, DiffRender.py:331/:360): Gauss-Bonnet exactly, 1/r^2 on spheres, sign on a torus."""
import math

import numpy as np
import torch

from conftest import uv_sphere


def _areas(V, F):
    tri = V[F]
    return 0.5 * torch.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0], dim=-1).norm(dim=1)


def _torus(R=0.6, r=0.2, nu=64, nv=32):
    u = np.linspace(0, 2 * np.pi, nu, endpoint=False)
    v = np.linspace(0, 2 * np.pi, nv, endpoint=False)
    uu, vv = np.meshgrid(u, v, indexing="ij")
    V = np.stack([(R + r * np.cos(vv)) * np.cos(uu), (R + r * np.cos(vv)) * np.sin(uu), r * np.sin(vv)], -1).reshape(-1, 3)
    idx = lambda i, j: (i % nu) * nv + (j % nv)
    F = []
    for i in range(nu):
        for j in range(nv):
            F += [[idx(i, j), idx(i + 1, j), idx(i + 1, j + 1)], [idx(i, j), idx(i + 1, j + 1), idx(i, j + 1)]]
    return V, np.asarray(F), vv.reshape(-1)


def test_gauss_bonnet_and_sphere_value():
    from nu_nerf_b200.tracer import discrete_gaussian_curvature
    for radius in (0.6, 1.5):
        V, F = uv_sphere(radius, 96, 48)
        V, F = torch.as_tensor(V, dtype=torch.float64), torch.as_tensor(F, dtype=torch.long)
        k = discrete_gaussian_curvature(V, F, clip=1e9).double().flatten()
        varea = torch.zeros(V.shape[0], dtype=torch.float64).index_add_(0, F.flatten(), (_areas(V, F) / 3.0).repeat_interleave(3))
        assert abs((k * varea).sum().item() - 4.0 * math.pi) < 1e-4            # sum of angle defects = 2 pi chi, chi = 2
        mid = (V[:, 2].abs() < 0.8 * radius)                                    # away from the poles of the UV grid
        assert (k[mid] * radius ** 2 - 1.0).abs().max().item() < 0.02
    assert discrete_gaussian_curvature(V, F).shape == (V.shape[0], 1)
    assert discrete_gaussian_curvature(torch.as_tensor(uv_sphere(0.1, 24, 12)[0]), torch.as_tensor(uv_sphere(0.1, 24, 12)[1])).max().item() <= 10.0


def test_torus_sign_and_zero_total():
    from nu_nerf_b200.tracer import discrete_gaussian_curvature
    V, F, vv = _torus()
    V, F = torch.as_tensor(V, dtype=torch.float64), torch.as_tensor(F, dtype=torch.long)
    k = discrete_gaussian_curvature(V, F, clip=1e9).double().flatten()
    varea = torch.zeros(V.shape[0], dtype=torch.float64).index_add_(0, F.flatten(), (_areas(V, F) / 3.0).repeat_interleave(3))
    assert abs((k * varea).sum().item()) < 1e-6                                  # chi = 0
    outer, inner = torch.from_numpy(np.cos(vv) > 0.5), torch.from_numpy(np.cos(vv) < -0.5)
    assert (k[outer] > 0).all() and (k[inner] < 0).all()
    # analytic: K = cos v / (r (R + r cos v))
    ref = torch.from_numpy(np.cos(vv) / (0.2 * (0.6 + 0.2 * np.cos(vv))))
    assert ((k - ref).abs() / ref.abs().clamp_min(1.0)).max().item() < 0.05

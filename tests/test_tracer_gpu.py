"""The two tracer OBJECTS of the boundary against the brute-force C oracle:
  optix_mesh   network/tracing_optix.py:119-158   update_mesh(F int32, V float32) / update_vert(V) / intersect(ray[N,6])
               -> (hit float32 in {1.0, 0.0}, idx int32 with 10000000 on a miss: cuda/triangle.cu:85-89)
  RayTracer    raytracing/raytracing/raytracer.py:8-54   RayTracer(vertices ndarray, triangles ndarray).trace(o, d, inplace)
               -> (positions, face normals, depth), depth clamped at MAX_DIST = 10 (raytracing/src/bvh.cu:36)
"""
import ctypes

import numpy as np
import pytest
import torch

from conftest import np_ptr, uv_sphere

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _oracle_hits(oracle_c, V, F, o, d, tmax=1e16):
    tv = np.ascontiguousarray(V[F].reshape(-1, 9).astype(np.float32))
    n = o.shape[0]
    h, ti, tt, uv = np.zeros(n, np.float32), np.zeros(n, np.int32), np.zeros(n, np.float32), np.zeros((n, 2), np.float32)
    oracle_c.oracle_closest_hit(np_ptr(tv), tv.shape[0], np_ptr(np.ascontiguousarray(o)), np_ptr(np.ascontiguousarray(d)),
                                n, ctypes.c_float(tmax), np_ptr(h), np_ptr(ti), np_ptr(tt), np_ptr(uv))
    return h, ti, tt


def _rays(n, seed=11):
    from oracle import nunerf_oracle as orc
    o, d = orc.synthetic_rays(n, seed=seed)
    o[::3] = 0.1 * o[::3]                 # a third of the rays start inside the mesh
    d[1::7] = -d[1::7]                    # some point away from it (misses)
    return o.contiguous(), d.contiguous()


def test_optix_mesh_update_and_intersect(oracle_c):
    from nu_nerf_b200.tracer import MISS_ID, TriangleBVH, optix_mesh
    V, F = uv_sphere(0.6, 48, 24)
    Vf, Fi = V.astype(np.float32), F.astype(np.int32)
    om = optix_mesh()
    with pytest.raises(RuntimeError):
        om.intersect(torch.zeros(4, 6, device=DEV))              # no mesh yet
    om.update_mesh(torch.from_numpy(Fi).to(DEV), torch.from_numpy(Vf).to(DEV))
    o, d = _rays(3000)
    ray = torch.cat([o, d], 1).to(DEV)
    hit, idx = om.intersect(ray)
    assert hit.dtype == torch.float32 and idx.dtype == torch.int32 and hit.shape == idx.shape == (3000,)
    h, ti, _ = _oracle_hits(oracle_c, Vf, F, o.numpy(), d.numpy())
    assert np.array_equal(idx.cpu().numpy(), ti), "triangle ids must be bit exact"
    assert np.array_equal(hit.cpu().numpy(), h)
    assert set(np.unique(hit.cpu().numpy())) <= {0.0, 1.0}
    miss = hit.cpu().numpy() == 0
    assert miss.any() and (idx.cpu().numpy()[miss] == MISS_ID).all()
    # the reference reshapes whatever it is given to [-1, 6] rows (tracing_optix.py:155)
    hit2, idx2 = om.intersect(ray.reshape(30, 100, 6))
    assert torch.equal(idx2, idx) and torch.equal(hit2, hit)
    # update_vert: same topology, moved vertices (DiffRender.py:404) -- the BVH is rebuilt on the new positions
    V2 = (Vf * np.array([1.3, 0.8, 1.0], np.float32) + np.array([0.05, 0.0, -0.02], np.float32)).astype(np.float32)
    om.update_vert(torch.from_numpy(V2).to(DEV))
    hit3, idx3 = om.intersect(ray)
    h3, ti3, _ = _oracle_hits(oracle_c, V2, F, o.numpy(), d.numpy())
    assert np.array_equal(idx3.cpu().numpy(), ti3) and np.array_equal(hit3.cpu().numpy(), h3)
    assert not np.array_equal(ti3, ti)
    assert TriangleBVH.overflow_count() == 0


def test_raytracer_trace_positions_normals_depth(oracle_c):
    from nu_nerf_b200.tracer import RayTracer
    V, F = uv_sphere(0.6, 32, 16)
    Vf = V.astype(np.float32)
    rt = RayTracer(Vf, F.astype(np.int32))
    with pytest.raises(AssertionError):
        RayTracer(Vf[:5], F[:8].astype(np.int32) % 5)           # "BVH needs at least 8 triangles!"
    o, d = _rays(2048, seed=12)
    o[5::11] = 20.0 * torch.nn.functional.normalize(o[5::11], dim=-1)      # farther than MAX_DIST: must read as a miss
    d[5::11] = -torch.nn.functional.normalize(o[5::11], dim=-1)
    pos, nrm, depth = rt.trace(o.to(DEV).view(32, 64, 3), d.to(DEV).view(32, 64, 3))
    assert pos.shape == nrm.shape == (32, 64, 3) and depth.shape == (32, 64)
    h, ti, tt = _oracle_hits(oracle_c, Vf, F, o.numpy(), d.numpy(), tmax=10.0)
    depth_ref = np.where(h > 0, tt, 10.0).astype(np.float32)
    assert np.array_equal(depth.reshape(-1).cpu().numpy(), depth_ref), "depth: closest t, clamped at MAX_DIST = 10"
    assert (depth_ref[5::11] == 10.0).all() and (h[5::11] == 0).all()
    pos_ref = o.numpy() + depth_ref[:, None] * d.numpy()
    assert np.abs(pos.reshape(-1, 3).cpu().numpy() - pos_ref).max() < 1e-5
    tri = Vf[F]
    fn = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    fn /= np.linalg.norm(fn, axis=1, keepdims=True)
    n_ref = np.where((h > 0)[:, None], fn[np.minimum(ti, F.shape[0] - 1)], 0.0)
    assert np.abs(nrm.reshape(-1, 3).cpu().numpy() - n_ref).max() < 1e-5
    # inplace=True overwrites the ray buffers with positions / normals (raytracer.py:39-47)
    ob, db = o.to(DEV).clone(), d.to(DEV).clone()
    p2, n2, d2 = rt.trace(ob, db, inplace=True)
    assert p2.data_ptr() == ob.data_ptr() and n2.data_ptr() == db.data_ptr()
    assert torch.equal(ob, pos.reshape(-1, 3)) and torch.equal(db, nrm.reshape(-1, 3)) and torch.equal(d2, depth.reshape(-1))


def test_bvh_builder_refuses_what_the_traversal_cannot_hold():
    """The traversal stack can never overflow silently: the builder bounds the depth, and a dropped push is counted."""
    from nu_nerf_b200.tracer import TriangleBVH
    V, F = uv_sphere(0.6, 224, 224)          # 99 904 triangles (config 4)
    bvh = TriangleBVH(torch.from_numpy(V.astype(np.float32)).to(DEV), torch.from_numpy(F.astype(np.int32)).to(DEV))
    o, d = _rays(20000, seed=13)
    hit, _ = bvh.trace(o.to(DEV), d.to(DEV))
    assert hit.sum().item() > 5000
    assert TriangleBVH.overflow_count() == 0


def test_dintersect_interpolates_the_vertex_curvature():
    """Scene.Dintersect (DiffRender.py:539-549): u, v, t, interpolated normal AND the interpolated vertex Gaussian
    curvature g_k (DiffRender.py:113-116) -- on a sphere of radius r every hit sees 1 / r^2 and the unit normal x / r."""
    from nu_nerf_b200.tracer import Scene
    r = 0.6
    V, Fc = uv_sphere(r, 96, 48)
    sc = Scene(V, Fc, device=DEV)
    g = torch.Generator().manual_seed(3)
    o = 3.0 * torch.nn.functional.normalize(torch.randn(512, 3, generator=g), dim=-1)
    d = torch.nn.functional.normalize(-o + 0.1 * torch.randn(512, 3, generator=g), dim=-1)
    o[:, 2] *= 0.3                                               # stay away from the poles of the UV grid
    d = torch.nn.functional.normalize(-o + 0.05 * torch.randn(512, 3, generator=g), dim=-1)
    info, hit = sc.Dintersect(o.to(DEV), d.to(DEV))
    assert hit.float().mean().item() > 0.9
    gk = info["g_k"][hit].flatten()
    assert info["g_k"].shape == (512, 1) and (info["g_k"][~hit] == 0).all()
    assert (gk * r * r - 1.0).abs().max().item() < 0.03
    x, n = info["x"][hit], torch.nn.functional.normalize(info["n"][hit], dim=-1)
    assert (x.norm(dim=-1) - r).abs().max().item() < 2e-3 and (n - x / r).abs().max().item() < 5e-3

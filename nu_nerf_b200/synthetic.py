"""Seeded synthetic inputs of SURVEY 8(d) for benchmarks and tools (no dataset, no network access): rays on a radius-3
sphere looking at the origin with 0.3 jitter, the two uniform draws of sample_ray, random target colours.  Pure torch CPU
generators, so every rank / run draws the same values (the test oracle keeps its own copy of these three functions)."""
import math

import numpy as np
import torch
import torch.nn.functional as F


def synthetic_rays(R, seed=1):
    g = torch.Generator().manual_seed(seed)
    o = 3.0 * F.normalize(torch.randn(R, 3, generator=g), dim=-1)
    d = F.normalize(-o + 0.3 * torch.randn(R, 3, generator=g), dim=-1)
    return o, d


def synthetic_uniforms(R, seed=2):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(R, 1, generator=g), torch.rand(R, 32, generator=g)


def synthetic_targets(R, seed=3):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(R, 3, generator=g)


def uv_sphere(radius=0.6, nu=48, nv=24):
    """Closed UV sphere (the synthetic outer mesh; the same construction generated the stage-2
    goldens).  float64 vertices [V,3], int64 faces [F,3], counter-clockwise seen from outside."""
    import math
    verts = [[0.0, 0.0, radius]]
    for i in range(1, nv):
        th = math.pi * i / nv
        for j in range(nu):
            ph = 2.0 * math.pi * j / nu
            verts.append([radius * math.sin(th) * math.cos(ph), radius * math.sin(th) * math.sin(ph),
                          radius * math.cos(th)])
    verts.append([0.0, 0.0, -radius])
    faces = []
    ring = lambda i, j: 1 + (i - 1) * nu + (j % nu)
    for j in range(nu):
        faces.append([0, ring(1, j), ring(1, j + 1)])
    for i in range(1, nv - 1):
        for j in range(nu):
            a, b, c, d = ring(i, j), ring(i, j + 1), ring(i + 1, j), ring(i + 1, j + 1)
            faces.append([a, c, d])
            faces.append([a, d, b])
    last = len(verts) - 1
    for j in range(nu):
        faces.append([last, ring(nv - 1, j + 1), ring(nv - 1, j)])
    return np.asarray(verts, dtype=np.float64), np.asarray(faces, dtype=np.int64)


def torus(R=0.55, r=0.22, nu=40, nv=20):
    """Closed torus around the z axis (both signs of Gaussian curvature; the non-zero-thickness test mesh)."""
    verts, faces = [], []
    for i in range(nu):
        ph = 2.0 * math.pi * i / nu
        for j in range(nv):
            th = 2.0 * math.pi * j / nv
            verts.append([(R + r * math.cos(th)) * math.cos(ph), (R + r * math.cos(th)) * math.sin(ph), r * math.sin(th)])
    idx = lambda i, j: (i % nu) * nv + (j % nv)
    for i in range(nu):
        for j in range(nv):
            a, b, c, d = idx(i, j), idx(i + 1, j), idx(i + 1, j + 1), idx(i, j + 1)
            faces.append([a, b, c])
            faces.append([a, c, d])
    return np.asarray(verts, dtype=np.float64), np.asarray(faces, dtype=np.int64)


def make_stage2(precision="split", mesh=None, thick=False, sphere_direction=False):
    """Stage2Renderer on the synthetic nested-sphere scene, built the way tests/golden/make_golden_stage2.py builds the
    reference's: stage-1 checkpoint
    from a seed-0 random-init NeROShapeRenderer, stage-2 modules from seed 5, in-memory UV-sphere outer mesh."""
    from nu_nerf_b200.renderer_zerothick import NeROShapeRenderer, load_default_cfg, name2renderer
    torch.manual_seed(0)
    cfg1 = load_default_cfg()
    cfg1["precision"] = precision
    if sphere_direction:            # the shader variant of the reference's real-data configs (configs/shape/real/*.yaml)
        cfg1["shader_config"] = {"sphere_direction": True, "human_light": False}
    net1 = NeROShapeRenderer(cfg1, training=False)
    cfg = {"name": "spherepot_s2", "network": "stage2", "database_name": "nerf/spherepot",
           "shader_config": {"sphere_direction": bool(sphere_direction), "human_light": False}, "apply_occ_loss": True,
           "occ_loss_step": 20000, "is_nerf": True, "zero_thickness": True, "eikonal_weight": 0.02,
           "freeze_inv_s_step": 5000, "precision": precision,
           "stage1_ckpt_dir": {"network_state_dict": net1.state_dict()}, "stage1_cfg_dir": cfg1,
           "stage1_mesh_dir": mesh if mesh is not None else uv_sphere()}
    torch.manual_seed(5)
    if thick:                       # network/renderer.py: the non-zero-thickness renderer (glass shell of learned thickness)
        from nu_nerf_b200.renderer import name2renderer as nz
        cfg["zero_thickness"], cfg["get_mask"] = False, False
        return nz["stage2"](cfg, training=False)
    return name2renderer["stage2"](cfg, training=False)

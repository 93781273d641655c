"""TEST INFRASTRUCTURE ONLY -- import shim for the unmodified reference at /root/reference.

Used in the build container (where /root/reference exists) to
  * pin the oracle restatement (oracle/nunerf_oracle.py, oracle/sampling_oracle.c) against
    outputs of the reference's own code, and
  * generate the golden fixtures committed under tests/golden/ (tests/golden/make_golden.py).

It never runs on the GPU box (no /root/reference there) and nothing in the product package
imports it.  The shims below are the deviations listed in SURVEY.md section 8(c):

  1. sys.modules stubs for third-party imports that are absent here (open3d, trimesh, pymesh, ...).
  2. np.math = math (utils/ref_utils.py:9,26-35 use np.math.factorial, removed in numpy 2).
  3. CUDA placement made a no-op on this CPU-only container (Tensor.cuda, device='cuda*' kwargs,
     set_default_tensor_type('torch.cuda.FloatTensor')).
  4. nvdiffrast dr.texture(tex[1,H,W,C], uv[1,P,1,2], 'linear', 'clamp') -> bilinear with texel
     centres at (i+0.5)/N and clamp-to-edge (F.grid_sample, align_corners=False, border).
     Parity unpinned: nvdiffrast is not vendored in the reference (field.py:721).
  5. optix_mesh.intersect -> brute force closest hit (oracle definition, see tracing section).
  6. torch.rand / torch.randperm injection so both sides consume identical uniform draws.
  7. (non-zero-thickness module, network/renderer.py: load_stage1 / load_stage2 with thick=True) the vertex Gaussian
     curvature DiffRender.Scene reads from PyMesh (DiffRender.py:331, :360) is angle_defect_curvature() below -- a stated
     definition (angle defect / barycentric vertex area, clipped to [-10, 10]); parity against PyMesh itself unpinned.
"""
import contextlib
import math
import os
import sys
import types
from unittest import mock

import numpy as np
import torch
import torch.nn.functional as F

REF_ROOT = os.environ.get("NUNERF_REFERENCE_ROOT", "/root/reference")

_STUBS = [
    "open3d", "trimesh", "trimesh.exchange", "trimesh.exchange.export", "trimesh.curvature", "pymesh",
    "skimage", "skimage.io", "skimage.metrics", "skimage.transform", "h5py", "plyfile", "transforms3d",
    "transforms3d.axangles", "transforms3d.euler", "transforms3d.quaternions", "imageio", "optix", "cupy",
    "mcubes", "nvdiffrast", "nvdiffrast.torch", "matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.lines",
    "matplotlib.backends", "matplotlib.backends.backend_agg", "matplotlib.figure",
    "tensorboardX", "pymeshlab", "tqdm", "lpips", "kornia", "pyexr", "OpenEXR", "Imath",
]


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "network"))


def _dr_texture(tex, uv, filter_mode="linear", boundary_mode="clamp"):
    # tex [1,H,W,C]; uv [1,P,1,2] with uv[...,0] -> width axis, uv[...,1] -> height axis.
    assert filter_mode == "linear" and boundary_mode == "clamp"
    out = F.grid_sample(tex.permute(0, 3, 1, 2), uv * 2.0 - 1.0, mode="bilinear",
                        padding_mode="border", align_corners=False)
    return out.permute(0, 2, 3, 1)  # [1,P,1,C]


def _strip_device(fn):
    def wrapped(*a, **k):
        dev = k.get("device", None)
        if dev is not None and "cuda" in str(dev):
            k.pop("device")
        return fn(*a, **k)
    return wrapped


_installed = False


def install():
    """Install the shims (idempotent) and put the reference on sys.path."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    for name in _STUBS:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = mock.MagicMock(name=name)
    dr = types.ModuleType("nvdiffrast.torch")
    dr.texture = _dr_texture
    sys.modules["nvdiffrast.torch"] = dr
    sys.modules["nvdiffrast"].torch = dr
    import importlib.machinery
    dr.__spec__ = importlib.machinery.ModuleSpec("nvdiffrast.torch", None)
    tq = types.ModuleType("tqdm")
    tq.__spec__ = importlib.machinery.ModuleSpec("tqdm", None)     # torch._dynamo walks sys.modules with find_spec
    tq.tqdm = lambda x, *a, **k: x
    tq.trange = lambda *a, **k: range(*a)
    sys.modules["tqdm"] = tq
    if not hasattr(np, "math"):
        np.math = math
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    for name in ["zeros", "ones", "zeros_like", "ones_like", "randperm", "linspace", "full", "empty", "arange",
                 "tensor", "eye", "rand", "randn"]:
        setattr(torch, name, _strip_device(getattr(torch, name)))
    _orig_set = torch.set_default_tensor_type
    torch.set_default_tensor_type = lambda t: _orig_set("torch.FloatTensor")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    _installed = True


@contextlib.contextmanager
def in_ref_dir():
    """The reference opens 'assets/bsdf_256_256.bin' relative to the cwd (field.py:583)."""
    cwd = os.getcwd()
    os.chdir(REF_ROOT)
    try:
        yield
    finally:
        os.chdir(cwd)


@contextlib.contextmanager
def injected_rand(draws):
    """Make torch.rand return the pre-drawn tensors in `draws` in call order (ZT:585,591)."""
    it = iter(draws)
    orig = torch.rand

    def fake(*a, **k):
        t = next(it)
        shape = list(a[0]) if len(a) == 1 and isinstance(a[0], (list, tuple, torch.Size)) else list(a)
        assert list(t.shape) == shape, (t.shape, shape)
        return t.clone()
    torch.rand = fake
    try:
        yield
    finally:
        torch.rand = orig


def load_stage1(seed=0, cfg_overrides=None, fg_lut=None, thick=False):
    """NeROShapeRenderer(cfg, training=False) of renderer_zerothick (or, thick=True, of the non-zero-thickness
    network/renderer.py) with spherepot.yaml."""
    install()
    with in_ref_dir():
        from utils.base_utils import load_cfg
        if thick:
            from network.renderer import NeROShapeRenderer
        else:
            from network.renderer_zerothick import NeROShapeRenderer
        cfg = load_cfg("configs/shape/nerf/spherepot.yaml")
        cfg.update(cfg_overrides or {})
        torch.manual_seed(seed)
        net = NeROShapeRenderer(cfg, training=False)
    if fg_lut is not None:
        net.color_network.FG_LUT.copy_(torch.as_tensor(fg_lut).reshape(1, 256, 256, 2))
    return net, cfg


def synthetic_rays(R, seed=1):
    """o = 3*normalize(randn), d = normalize(-o + 0.3*randn)."""
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


# ------------------------------------------------------------------------------------------------ stage 2
def uv_sphere(radius=0.6, nu=48, nv=24):
    """Closed UV sphere (the synthetic outer mesh): nu longitudes, nv latitude bands.
    Returns float64 vertices [V,3] and int64 faces [F,3], counter-clockwise seen from outside."""
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


def brute_force_closest_hit(V, Fc, o, d, chunk=256):
    """Oracle definition of optix_mesh.intersect (SURVEY 8c shim 6): double-sided Moeller-Trumbore in fp32 in the
    operation order of JIT_Dintersect, 0 < t < 1e16, u,v >= 0, u+v <= 1; closest hit, ties: min t then min face id.
    Returns (hit float32 {0,1}, idx int32 with 10000000 on a miss)."""
    tri = V[Fc.long()].float()
    v0, e1, e2 = tri[:, 0], tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]
    N = o.shape[0]
    hit = torch.zeros(N)
    idx = torch.full((N,), 10000000, dtype=torch.int32)
    for s in range(0, N, chunk):
        oo, dd = o[s:s + chunk].float()[:, None, :], d[s:s + chunk].float()[:, None, :]
        pvec = torch.cross(dd.expand(-1, e2.shape[0], -1), e2[None].expand(oo.shape[0], -1, -1), dim=-1)
        det = (e1[None] * pvec).sum(-1)
        inv = 1.0 / det
        tvec = oo - v0[None]
        u = (tvec * pvec).sum(-1) * inv
        qvec = torch.cross(tvec, e1[None].expand(oo.shape[0], -1, -1), dim=-1)
        v = (dd * qvec).sum(-1) * inv
        t = (e2[None] * qvec).sum(-1) * inv
        ok = (det != 0) & (u >= 0) & (v >= 0) & (u + v <= 1) & (t > 0) & (t < 1e16)
        tt = torch.where(ok, t, torch.full_like(t, float("inf")))
        tmin, _ = tt.min(dim=1)
        first = (tt == tmin[:, None]) & ok
        fid = torch.where(first, torch.arange(tt.shape[1])[None], torch.full_like(tt, 1 << 30, dtype=torch.long)).min(1)[0]
        h = torch.isfinite(tmin)
        hit[s:s + chunk] = h.float()
        idx[s:s + chunk] = torch.where(h, fid, torch.full_like(fid, 10000000)).int()
    return hit, idx


def angle_defect_curvature(V, Fc, clip=10.0):
    """The stated replacement for PyMesh's vertex_gaussian_curvature (DiffRender.py:331, :360; clipped to [-10, 10] there),
    restated in numpy for the oracle side: (2 pi - sum of corner angles) / (one third of the incident triangle area).
    Parity against PyMesh itself is unpinned (the fork the reference builds against is not available)."""
    V = np.asarray(V, np.float64)
    Fc = np.asarray(Fc, np.int64)
    tri = V[Fc]
    area = 0.5 * np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1)
    defect = np.full(V.shape[0], 2.0 * np.pi)
    varea = np.zeros(V.shape[0])
    for i in range(3):
        a = tri[:, (i + 1) % 3] - tri[:, i]
        b = tri[:, (i + 2) % 3] - tri[:, i]
        cosv = (a * b).sum(-1) / (np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1))
        np.add.at(defect, Fc[:, i], -np.arccos(np.clip(cosv, -1.0, 1.0)))
        np.add.at(varea, Fc[:, i], area / 3.0)
    return np.clip(defect / np.maximum(varea, 1e-30), -clip, clip).astype(np.float32).reshape(-1, 1)


def torus(R=0.55, r=0.22, nu=40, nv=20):
    """Closed torus around the z axis (both curvature signs: the non-zero-thickness shell offset has a branch per sign).
    float64 vertices [nu*nv,3], int64 faces [2*nu*nv,3], counter-clockwise seen from outside."""
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


def load_stage2(mesh_V, mesh_F, stage1_seed=0, seed=5, fg_lut=None, tmp_dir="/tmp", thick=False, sphere_direction=False):
    """Stage2Renderer(cfg, training=False) of renderer_zerothick (thick=False) or of the non-zero-thickness
    network/renderer.py (thick=True) with configs/stage2/nerf/spherepot.yaml, a stage-1
    checkpoint made from a seeded random-init NeROShapeRenderer of the same module, and the in-memory mesh (shim 5: Scene
    built from V, F with the reference's own corner-angle vertex normals; vertex Gaussian curvature = zero for the
    zero-thickness path, which never reads it, and angle_defect_curvature() for the non-zero-thickness one)."""
    install()
    with in_ref_dir():
        from utils.base_utils import load_cfg
        import network.DiffRender as DR
        if thick:
            import network.renderer as ZT
        else:
            import network.renderer_zerothick as ZT
        cfg1 = load_cfg("configs/shape/nerf/spherepot.yaml")
        if sphere_direction:        # the shader variant of the real-data configs (configs/shape/real/*.yaml)
            cfg1 = {**cfg1, "shader_config": {**cfg1.get("shader_config", {}), "sphere_direction": True}}
            cfg1_path = os.path.join(tmp_dir, "nunerf_stage1_cfg_sph.yaml")
            import yaml
            with open(cfg1_path, "w") as f:
                yaml.safe_dump(cfg1, f)
        torch.manual_seed(stage1_seed)
        net1 = ZT.NeROShapeRenderer(cfg1, training=False)
    if fg_lut is not None:
        net1.color_network.FG_LUT.copy_(torch.as_tensor(fg_lut).reshape(1, 256, 256, 2))
    ckpt = os.path.join(tmp_dir, ("nunerf_stage1_ckpt_nz%s.pth" % ("_sph" if sphere_direction else "")) if thick
                        else "nunerf_stage1_ckpt.pth")
    torch.save({"network_state_dict": net1.state_dict(), "step": 0}, ckpt)
    curv = angle_defect_curvature(mesh_V, mesh_F) if thick else np.zeros((len(mesh_V), 1), np.float32)
    with in_ref_dir():

        class ShimOptix:
            def update_mesh(self, F_, V_):
                self.F, self.V = F_.long(), V_.float()

            def update_vert(self, V_):
                self.V = V_.float()

            def intersect(self, ray):
                return brute_force_closest_hit(self.V, self.F, ray[:, :3], ray[:, 3:])

        class ShimScene(DR.Scene):
            def __init__(self, mesh_path, cuda_device=0):
                self.optix_mesh = ShimOptix()
                self.vertices = torch.tensor(mesh_V, dtype=torch.float64)
                self.faces = torch.tensor(mesh_F, dtype=torch.long)
                self.triangles = self.vertices[self.faces]
                self.optix_mesh.update_mesh(self.faces.to(torch.int32), self.vertices.to(torch.float32))
                # init_VN (DiffRender.py:342-359) with the reference's JIT_corner_angles
                corner_angles, face_N = DR.JIT_corner_angles(self.triangles)
                row = self.faces.view(-1)
                col = torch.arange(len(self.faces)).unsqueeze(1).expand(-1, 3).reshape(-1)
                M = torch.sparse_coo_tensor(torch.stack((row, col)), corner_angles.detach(),
                                            (len(self.vertices), len(self.faces)))
                vert_N = torch.sparse.mm(M, face_N)
                self.normals = vert_N / vert_N.norm(dim=1, p=2, keepdim=True)
                self.gaussian_curvatures = torch.tensor(curv).reshape(-1, 1)          # DiffRender.py:360

        DR.device = "cpu"            # module-level placement constant of DiffRender.py:16 (shim 3)
        ZT.Scene = ShimScene
        cfg = load_cfg("configs/stage2/nerf/spherepot.yaml")
        cfg["stage1_ckpt_dir"] = ckpt
        cfg["stage1_cfg_dir"] = cfg1_path if sphere_direction else "configs/shape/nerf/spherepot.yaml"
        if sphere_direction:
            cfg["shader_config"] = {**cfg.get("shader_config", {}), "sphere_direction": True}
        cfg["stage1_mesh_dir"] = "<in-memory>"
        torch.manual_seed(seed)
        net = ZT.Stage2Renderer(cfg, training=False)
    if fg_lut is not None:
        lut = torch.as_tensor(fg_lut).reshape(1, 256, 256, 2)
        net.stage1_network.color_network.FG_LUT.copy_(lut)
        net.color_network_inner.FG_LUT.copy_(lut)
    return net, cfg

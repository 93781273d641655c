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
    "mcubes", "nvdiffrast", "nvdiffrast.torch", "matplotlib", "matplotlib.pyplot", "matplotlib.cm",
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
    tq = types.ModuleType("tqdm")
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


def load_stage1(seed=0, cfg_overrides=None, fg_lut=None):
    """NeROShapeRenderer(cfg, training=False) of renderer_zerothick with spherepot.yaml."""
    install()
    with in_ref_dir():
        from utils.base_utils import load_cfg
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

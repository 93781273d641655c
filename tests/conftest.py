import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def reference_fg_lut(tmp_path_factory):
    """Every renderer built by the tests registers the REFERENCE's split-sum table (tests/golden/fg_lut_reference.npz =
    its asset assets/bsdf_256_256.bin, which the goldens were generated with) through the product's own search order
    (nu_nerf_b200/fg_lut.load_fg_lut: $NUNERF_FG_LUT first)."""
    lut = np.load(os.path.join(GOLDEN, "fg_lut_reference.npz"))["FG_LUT"].astype(np.float32)
    path = str(tmp_path_factory.mktemp("lut") / "bsdf_256_256.bin")
    lut.tofile(path)
    old = os.environ.get("NUNERF_FG_LUT")
    os.environ["NUNERF_FG_LUT"] = path
    yield lut
    if old is None:
        os.environ.pop("NUNERF_FG_LUT", None)
    else:
        os.environ["NUNERF_FG_LUT"] = old


def _make(target):
    subprocess.run(["make", "-s", target], cwd=ROOT, check=True)


@pytest.fixture(scope="session")
def oracle_c():
    """The plain-C oracle (oracle/sampling_oracle.c), built on demand with gcc."""
    path = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(ROOT, "oracle", "sampling_oracle.c")):
        _make("oracle/_build/liboracle.so")
    return ctypes.CDLL(path)


@pytest.fixture(scope="session")
def stage1_sd():
    """state_dict of a freshly initialised stage-1 renderer (bit-identical to the reference's, see test_init)."""
    from nu_nerf_b200.renderer_zerothick import NeROShapeRenderer, load_default_cfg
    torch.manual_seed(0)
    net = NeROShapeRenderer(load_default_cfg(), training=False)
    return {k: v.detach().clone() for k, v in net.state_dict().items()}


def np_ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def sampling_tables():
    """The linspace tables the sampling kernels take (computed with torch exactly as ZT:580-590 does)."""
    t = torch.linspace(0.0, 1.0, 64)
    b = torch.linspace(1e-3, 1.0 - 1.0 / 33.0, 32)
    mids = 0.5 * (b[1:] + b[:-1])
    upper = torch.cat([mids, b[-1:]])
    lower = torch.cat([b[:1], mids])
    return torch.cat([t, lower, upper - lower, b]).contiguous()


def uv_sphere(radius=0.6, nu=48, nv=24):
    """Closed UV sphere (same construction as oracle/ref_harness.uv_sphere, which generated the stage-2 goldens)."""
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


def make_stage2(precision="split", mesh=None):
    """Product Stage2Renderer built the way tests/golden/make_golden_stage2.py builds the reference's: stage-1 checkpoint
    from a seed-0 random-init NeROShapeRenderer, stage-2 modules from seed 5, in-memory UV-sphere outer mesh."""
    from nu_nerf_b200.renderer_zerothick import NeROShapeRenderer, load_default_cfg, name2renderer
    torch.manual_seed(0)
    cfg1 = load_default_cfg()
    cfg1["precision"] = precision
    net1 = NeROShapeRenderer(cfg1, training=False)
    cfg = {"name": "spherepot_s2", "network": "stage2", "database_name": "nerf/spherepot",
           "shader_config": {"sphere_direction": False, "human_light": False}, "apply_occ_loss": True,
           "occ_loss_step": 20000, "is_nerf": True, "zero_thickness": True, "eikonal_weight": 0.02,
           "freeze_inv_s_step": 5000, "precision": precision,
           "stage1_ckpt_dir": {"network_state_dict": net1.state_dict()}, "stage1_cfg_dir": cfg1,
           "stage1_mesh_dir": mesh if mesh is not None else uv_sphere()}
    torch.manual_seed(5)
    return name2renderer["stage2"](cfg, training=False)

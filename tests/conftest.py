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


from nu_nerf_b200.synthetic import make_stage2, uv_sphere  # noqa: E402,F401  (the synthetic nested-sphere scene of config 4)


def stage2_rec_from_golden(G, device):
    """The trace record Stage2Renderer._replay_geometry takes (hit / pass index lists, triangle ids, per-sample
    parameters z), rebuilt from the reference's own ray_trace lists in tests/golden/stage2_R64.npz: z is recovered from the
    sampled points (z = <p - start, delta> / |delta|^2 in float64)."""
    T = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a)).to(device) if dt is None else \
        torch.from_numpy(np.ascontiguousarray(a)).to(device).to(dt)
    n = int(G["n_segments"])
    rec = {"bounces": [], "segments": []}
    for k in range(n):
        hit = G[f"trace_hit_{k}"].reshape(-1) > 0
        conv = G[f"converge_{k}"].reshape(-1)
        path = G[f"path_{k}"].astype(np.float64)
        start, dirs = path[:, 0].copy(), G[f"dir_{k}"].astype(np.float64)
        if k != 1:
            start[~hit] -= 0.1 * dirs[~hit]          # rays that leave the scene are sampled on [0.1, 64] (ZT:1764)
        end = start + dirs * 4.5
        end[hit] = path[hit, -1]
        delta = end - start
        m_idx = None
        if k != 1 and (~hit).any():
            delta[~hit] = dirs[~hit]
            m_idx = T(np.nonzero(~hit)[0])
        Z = ((path - start[:, None]) * delta[:, None]).sum(-1) / (delta * delta).sum(-1)[:, None]
        rec["bounces"].append(dict(hit_idx=T(np.nonzero(hit)[0]), ok_idx=T(np.nonzero(conv[hit])[0]), inside=(k % 2 == 1),
                                   tri=T(G[f"trace_tri_{k}"].reshape(-1)[hit], torch.long)))
        rec["segments"].append(dict(Z=T(Z, torch.float32), h_idx=T(np.nonzero(hit)[0]), m_idx=m_idx,
                                    start=T(start, torch.float32), dir=T(dirs, torch.float32)))
    return rec

"""Marching cubes (reference call site network/field.py:1312 -> PyMCubes, unpinned and absent: parity
unpinned, see oracle/mc_oracle.py).  CPU tests pin the generated case tables through surface properties; GPU tests
compare csrc/mcubes.cu with the numpy oracle bit for bit and check the indexed mesh the public call returns."""
import numpy as np
import pytest
import torch

from oracle import mc_oracle as mco


def _sphere_grid(n, r=0.6, centre=(0.0, 0.0, 0.0)):
    g = np.linspace(-1, 1, n)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    return (np.sqrt((X - centre[0]) ** 2 + (Y - centre[1]) ** 2 + (Z - centre[2]) ** 2) - r).astype(np.float32)


def _index(keys):
    uniq, inv = np.unique(keys.reshape(-1), return_inverse=True)
    return inv.reshape(-1, 3)


def test_case_tables_known_answers():
    assert mco.check_tables()


def test_oracle_sphere_is_a_closed_oriented_manifold():
    n = 20
    soup, keys = mco.marching_cubes_soup(_sphere_grid(n), 0.0)
    V, E, F, boundary, nonmanifold, inconsistent = mco.mesh_stats(_index(keys))
    assert (boundary, nonmanifold, inconsistent) == (0, 0, 0)
    assert V - E + F == 2
    nrm = np.cross(soup[:, 1] - soup[:, 0], soup[:, 2] - soup[:, 0])
    assert (np.einsum("ij,ij->i", nrm, soup.mean(1) - (n - 1) / 2) > 0).all()     # table winding: towards larger u
    # vertices sit on the linear-interpolation crossing: radius within the chord error of the grid step
    rad = np.linalg.norm(soup.reshape(-1, 3) / (n - 1) * 2 - 1, axis=1)
    assert np.abs(rad - 0.6).max() < 0.5 * (2 / (n - 1)) ** 2 / 0.6 + 1e-6


def test_oracle_random_field_is_watertight_inside_the_grid():
    """All 256 cases incl. the ambiguous faces: edges may be open only on the border of the grid."""
    rng = np.random.default_rng(0)
    n = 12
    u = rng.standard_normal((n, n, n)).astype(np.float32)
    soup, keys = mco.marching_cubes_soup(u, 0.1)
    tri = _index(keys)
    V, E, F, boundary, nonmanifold, inconsistent = mco.mesh_stats(tri)
    assert nonmanifold == 0 and inconsistent == 0 and F > 1000
    pos = np.zeros((tri.max() + 1, 3), dtype=np.float32)
    pos[tri.reshape(-1)] = soup.reshape(-1, 3)
    half = np.concatenate([tri[:, [0, 1]], tri[:, [1, 2]], tri[:, [2, 0]]])
    und, cnt = np.unique(np.sort(half, axis=1), axis=0, return_counts=True)
    for a, b in und[cnt == 1]:
        on_border = ((pos[a] == 0) | (pos[a] == n - 1)) & ((pos[b] == 0) | (pos[b] == n - 1))
        assert on_border.any()


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["sphere20", "random12", "offcentre33", "res2", "empty"])
def test_device_soup_bit_exact_vs_oracle(case):
    from nu_nerf_b200 import sweep
    from nu_nerf_b200._lib import call
    rng = np.random.default_rng(1)
    u, iso = {"sphere20": (_sphere_grid(20), 0.0), "random12": (rng.standard_normal((12, 12, 12)).astype(np.float32), 0.1),
              "offcentre33": (_sphere_grid(33, 0.45, (0.2, -0.1, 0.3)), 0.05),
              "res2": (rng.standard_normal((2, 2, 2)).astype(np.float32), 0.0),
              "empty": (np.ones((9, 9, 9), dtype=np.float32), 0.0)}[case]
    soup_ref, keys_ref = mco.marching_cubes_soup(u, iso)
    ud = torch.from_numpy(u).cuda()
    res = u.shape[0]
    tri_table, n_tris, edges, edge_axis = sweep._mc_tables(ud.device)
    from nu_nerf_b200 import _lib
    n_blocks = int(_lib.lib.nunerf_mc_blocks(res))
    counts = torch.empty(n_blocks, dtype=torch.int32, device="cuda")
    call("nunerf_mc_count", ud.data_ptr(), res, float(iso), n_tris.data_ptr(), counts.data_ptr())
    assert int(counts.sum()) == len(soup_ref)
    v, t = sweep.marching_cubes(ud, iso)
    if len(soup_ref) == 0:
        assert v.shape == (0, 3) and t.shape == (0, 3)
        return
    offsets = (torch.cumsum(counts, 0, dtype=torch.int64) - counts).contiguous()
    total = len(soup_ref)
    verts = torch.empty(3 * total, 3, dtype=torch.float32, device="cuda")
    keys = torch.empty(3 * total, dtype=torch.int64, device="cuda")
    call("nunerf_mc_emit", ud.data_ptr(), res, float(iso), tri_table.data_ptr(), tri_table.shape[1] // 3, n_tris.data_ptr(),
         edges.data_ptr(), edge_axis.data_ptr(), offsets.data_ptr(), verts.data_ptr(), keys.data_ptr())
    assert np.array_equal(verts.cpu().numpy().reshape(-1, 3, 3), soup_ref)          # bit-exact, same cell order
    assert np.array_equal(keys.cpu().numpy().reshape(-1, 3), keys_ref)
    # the public call: indexed, PyMCubes orientation (reversed), same vertex set
    assert t.shape == (total, 3) and v.dtype == np.float64 and t.dtype == np.int64
    assert np.array_equal(v[t[:, [0, 2, 1]]].astype(np.float32), soup_ref)


@pytest.mark.gpu
def test_marching_cubes_sphere_128_properties():
    from nu_nerf_b200.sweep import marching_cubes
    n = 128
    v, t = marching_cubes(torch.from_numpy(_sphere_grid(n)).cuda(), 0.0)
    V, E, F, boundary, nonmanifold, inconsistent = mco.mesh_stats(t)
    assert (boundary, nonmanifold, inconsistent) == (0, 0, 0) and V - E + F == 2 and V == len(v)
    p = v[t]
    nrm = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
    assert (np.einsum("ij,ij->i", nrm, p.mean(1) - (n - 1) / 2) < 0).all()      # PyMCubes convention: towards smaller u
    rad = np.linalg.norm(v / (n - 1) * 2 - 1, axis=1)
    assert np.abs(rad - 0.6).max() < 1e-4
    area = 0.5 * np.linalg.norm(nrm, axis=1).sum() * (2 / (n - 1)) ** 2
    assert abs(area - 4 * np.pi * 0.36) / (4 * np.pi * 0.36) < 2e-3


@pytest.mark.gpu
def test_marching_cubes_rows_longer_than_one_block():
    """res - 1 > 256: a cell row spans two z-chunks of a block (running offsets across chunks and rows); off-centre
    ellipsoid so that the surface crosses the chunk boundary obliquely."""
    from nu_nerf_b200.sweep import marching_cubes
    n = 300
    g = torch.linspace(-1, 1, n, device="cuda")
    X, Y, Z = torch.meshgrid(g, g, g, indexing="ij")
    u = torch.sqrt(((X - 0.1) / 0.7) ** 2 + ((Y + 0.05) / 0.5) ** 2 + ((Z - 0.2) / 0.6) ** 2) - 1.0
    v, t = marching_cubes(u.contiguous(), 0.0)
    V, E, F, boundary, nonmanifold, inconsistent = mco.mesh_stats(t)
    assert (boundary, nonmanifold, inconsistent) == (0, 0, 0) and V - E + F == 2 and V == len(v)
    # every vertex sits on a grid edge (two integer coordinates) and on the linearly interpolated surface
    frac = np.abs(v - np.round(v))
    assert ((frac > 1e-6).sum(axis=1) <= 1).all()
    w = v / (n - 1) * 2 - 1
    f = np.sqrt(((w[:, 0] - 0.1) / 0.7) ** 2 + ((w[:, 1] + 0.05) / 0.5) ** 2 + ((w[:, 2] - 0.2) / 0.6) ** 2) - 1.0
    assert np.abs(f).max() < 2e-4


@pytest.mark.gpu
def test_extract_geometry_of_the_initial_field():
    """extract_geometry (field.py:1310-1319) end to end on the geometric-init field (an approximate sphere, radius 0.5)."""
    from nu_nerf_b200.renderer_zerothick import NeROShapeRenderer, load_default_cfg
    from nu_nerf_b200.sweep import extract_geometry
    torch.manual_seed(0)
    cfg = load_default_cfg()
    cfg["precision"] = "bf16"
    net = NeROShapeRenderer(cfg, training=False).cuda()
    v, t = extract_geometry(-torch.ones(3), torch.ones(3), 96, 0, net.sdf_network.sdf)
    V, E, F, boundary, nonmanifold, inconsistent = mco.mesh_stats(t)
    assert (boundary, nonmanifold, inconsistent) == (0, 0, 0) and V - E + F == 2
    rad = np.linalg.norm(v, axis=1)
    assert 0.3 < rad.min() and rad.max() < 0.7          # the geometric init is a bumpy sphere of radius ~0.5

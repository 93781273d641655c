"""TEST INFRASTRUCTURE ONLY (imported by tests/ only) -- CPU marching cubes used as the checker of csrc/mcubes.cu.

The reference calls PyMCubes (`mcubes.marching_cubes(u, threshold)`, network/field.py:1312).  PyMCubes is an un-vendored,
unpinned dependency (requirements.txt:14) and is absent here, so this oracle restates the published algorithm (Lorensen &
Cline: classify the 8 corners of every cell against the iso-value, look the case up in a triangle table, place one vertex
per crossed edge by linear interpolation).  **Parity unpinned** against PyMCubes itself: its table may triangulate the
polygons of a cell differently.  What is pinned are properties every correct table gives (tests/test_mcubes_*.py):
closed 2-manifold output with consistent orientation on closed surfaces, Euler characteristic, vertices exactly on the
linearly interpolated crossings, orientation convention of PyMCubes (normals towards decreasing u, which is why
extract_mesh_stage1.py:41 flips the faces).

`marching_cubes_soup` walks the cells in x-major order with fp32 arithmetic in the order the kernel uses, so the CUDA
triangle soup can be compared bit for bit.  The CLASSIFICATION and INTERPOLATION here are independent code; the case
table is the one the product generates (nu_nerf_b200.mc_tables) and is checked on its own by `check_tables`.
"""
import numpy as np

from nu_nerf_b200.mc_tables import CORNERS, build_tables


def marching_cubes_soup(u, iso):
    """u [n,n,n] float32 -> triangle soup float32 [T,3,3] (grid-index coordinates; winding of the generated table =
    normals towards larger u) and int64 edge keys [T,3]."""
    table, n_tris, edges, axis = [t.astype(np.int64) for t in build_tables()]
    u = np.asarray(u, dtype=np.float32)
    iso = np.float32(iso)
    n = u.shape[0]
    soup, keys = [], []
    for x in range(n - 1):
        for y in range(n - 1):
            for z in range(n - 1):
                vals = [u[x + CORNERS[i][0], y + CORNERS[i][1], z + CORNERS[i][2]] for i in range(8)]
                case = sum(1 << i for i in range(8) if vals[i] < iso)
                for t in range(n_tris[case]):
                    for e in table[case, 3 * t:3 * t + 3]:
                        c0, c1 = edges[e]
                        a = np.float32(iso - vals[c0]) / np.float32(vals[c1] - vals[c0])
                        p0 = np.array([x, y, z]) + CORNERS[c0]
                        p = p0.astype(np.float32)
                        p[axis[e]] = np.float32(p0[axis[e]]) + np.float32(a)
                        soup.append(p)
                        keys.append(((p0[0] * n + p0[1]) * n + p0[2]) * 3 + axis[e])
    return np.array(soup, dtype=np.float32).reshape(-1, 3, 3), np.array(keys, dtype=np.int64).reshape(-1, 3)


def mesh_stats(triangles):
    """(V, E, F, boundary_edges, non_manifold_edges, inconsistent_edges) of an indexed triangle list."""
    tri = np.asarray(triangles, dtype=np.int64)
    half = np.concatenate([tri[:, [0, 1]], tri[:, [1, 2]], tri[:, [2, 0]]])
    und = np.sort(half, axis=1)
    uniq, inv, cnt = np.unique(und, axis=0, return_inverse=True, return_counts=True)
    inv = inv.reshape(-1)
    # orientation: the two half-edges of an interior edge must run in opposite directions
    direction = np.where(half[:, 0] < half[:, 1], 1, -1)
    net = np.zeros(len(uniq), dtype=np.int64)
    np.add.at(net, inv, direction)
    return (len(np.unique(tri)), len(uniq), len(tri), int((cnt == 1).sum()), int((cnt > 2).sum()),
            int(((cnt == 2) & (net != 0)).sum()))


def check_tables():
    """Table-level known answers: complementary cases have the same triangle count (the ambiguous-face rule is
    sign-symmetric up to which corners are cut off, so only the NUMBER may differ by the rule; the vertex SET must be the
    crossed edges exactly), every triangle uses only crossed edges, and every crossed edge is used."""
    table, n_tris, edges, _ = build_tables()
    for case in range(256):
        crossed = {e for e, (c0, c1) in enumerate(edges.tolist()) if ((case >> c0) & 1) != ((case >> c1) & 1)}
        used = set(int(e) for e in table[case, :3 * n_tris[case]])
        assert used == crossed, (case, used, crossed)
        assert (table[case, 3 * n_tris[case]:] == -1).all()
    assert n_tris[0] == 0 and n_tris[255] == 0
    assert all(n_tris[1 << i] == 1 for i in range(8))
    return True

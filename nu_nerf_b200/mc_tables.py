"""Marching-cubes case tables, GENERATED (not transcribed): for each of the 256 corner-sign patterns the iso-surface
polygons of the cell are found by walking the crossed edges face by face, and fan-triangulated.

Conventions: corner i sits at offset (i & 1, i >> 1 & 1, i >> 2 & 1) along (x, y, z); bit i of the case index is set
when the corner value is BELOW the iso-value ("inside"); edge e joins EDGES[e] = (c0, c1), c0 < c1, along axis
EDGE_AXIS[e].  On a face with four crossed edges (two inside corners on a diagonal) the segments always cut off the
INSIDE corners; the rule depends on the face's own four signs only, so the two cells sharing a face agree and the mesh is
watertight.  Triangles are wound counter-clockwise seen from the outside (normals towards larger values).

This replaces the table of PyMCubes (`mcubes.marching_cubes`, field.py:1312; the package is not pinned by the reference
and absent here): same surface, possibly different but equally valid triangulation of the ambiguous cases.
"""
import numpy as np

CORNERS = np.array([[i & 1, (i >> 1) & 1, (i >> 2) & 1] for i in range(8)], dtype=np.int64)
EDGES = [(c, c | (1 << a)) for a in range(3) for c in range(8) if not (c >> a) & 1]
EDGE_AXIS = [a for a in range(3) for c in range(8) if not (c >> a) & 1]
_EDGE_ID = {e: i for i, e in enumerate(EDGES)}


def _faces():
    faces = []
    for a in range(3):
        b, c = [x for x in range(3) if x != a]
        for side in (0, 1):
            cyc = []
            for (vb, vc) in ((0, 0), (1, 0), (1, 1), (0, 1)):
                cyc.append((side << a) | (vb << b) | (vc << c))
            faces.append(cyc)
    return faces


FACES = _faces()


def _case_loops(case):
    inside = [(case >> i) & 1 for i in range(8)]
    nbr = {}

    def link(e0, e1):
        nbr.setdefault(e0, []).append(e1)
        nbr.setdefault(e1, []).append(e0)
    for cyc in FACES:
        fe = []          # (edge id, index of its first corner in the cycle) for the crossed edges of this face
        for k in range(4):
            c0, c1 = cyc[k], cyc[(k + 1) % 4]
            if inside[c0] != inside[c1]:
                fe.append((_EDGE_ID[(min(c0, c1), max(c0, c1))], k))
        if len(fe) == 2:
            link(fe[0][0], fe[1][0])
        elif len(fe) == 4:
            # ambiguous face: cut off each inside corner (edges k-1 and k meet at corner cyc[k])
            by_k = {k: e for e, k in fe}
            for k in range(4):
                if inside[cyc[k]]:
                    link(by_k[(k - 1) % 4], by_k[k])
    loops, seen = [], set()
    for start in sorted(nbr):
        if start in seen:
            continue
        loop, prev, cur = [start], None, start
        seen.add(start)
        while True:
            a, b = nbr[cur]
            nxt = a if a != prev else b
            if nxt == start and len(loop) > 2:
                break
            if nxt in seen:          # both neighbours already visited: closed
                break
            loop.append(nxt)
            seen.add(nxt)
            prev, cur = cur, nxt
        loops.append(loop)
    # orientation: Newell normal of the loop (edge midpoints) must point from the inside corners to the outside ones
    out = []
    for loop in loops:
        mid = np.array([(CORNERS[EDGES[e][0]] + CORNERS[EDGES[e][1]]) * 0.5 for e in loop])
        n = np.zeros(3)
        for i in range(len(loop)):
            p, q = mid[i], mid[(i + 1) % len(loop)]
            n += np.cross(p, q)
        ins = np.mean([CORNERS[c] for e in loop for c in EDGES[e] if inside[c]], axis=0)
        outs = np.mean([CORNERS[c] for e in loop for c in EDGES[e] if not inside[c]], axis=0)
        if np.dot(n, outs - ins) < 0:
            loop = loop[::-1]
        out.append(loop)
    return out


def build_tables():
    """(tri_table int8 [256, 3 * MAX_TRIS] padded with -1, n_tris int8 [256], edge_corners int8 [12, 2], edge_axis int8 [12])."""
    tris = []
    for case in range(256):
        t = []
        for loop in _case_loops(case):
            for i in range(1, len(loop) - 1):
                t += [loop[0], loop[i], loop[i + 1]]
        tris.append(t)
    max_tris = max(len(t) for t in tris) // 3
    table = -np.ones((256, 3 * max_tris), dtype=np.int8)
    for case, t in enumerate(tris):
        table[case, :len(t)] = t
    n_tris = np.array([len(t) // 3 for t in tris], dtype=np.int8)
    return table, n_tris, np.array(EDGES, dtype=np.int8), np.array(EDGE_AXIS, dtype=np.int8)

"""SDF grid sweep for mesh extraction: extract_fields / extract_geometry of network/field.py:1286-1319.

The reference evaluates the N^3 grid in 64^3 blocks, each with its own meshgrid, .cuda() upload, MLP pass and
.cpu() download (512 blocks and 512 host syncs at N = 512).  Here the grid coordinates are generated on the device,
the SDF comes from the fused inference kernel (csrc/chain.cu), outside-sphere masking is one more kernel, the whole
field stays in ONE contiguous device buffer and is copied to the host once.
"""
import numpy as np
import torch

from . import _lib
from ._lib import call


def _owning_renderer(query_func):
    net = getattr(query_func, "__self__", None)
    q = getattr(net, "_query", None)
    return getattr(q, "__self__", None)


@torch.no_grad()
def extract_fields(bound_min, bound_max, resolution, query_func, batch_size=64, outside_val=1.0, chunk_points=1 << 22,
                   return_device=False):
    """Same contract as the reference: returns u[res,res,res] float32 (numpy).  `query_func` is normally
    `renderer.sdf_network.sdf` (extract_mesh_stage1.py:36): then the sweep runs entirely on the device; any other
    callable [P,3] -> [P,1] is evaluated chunk by chunk (`batch_size` is accepted for signature compatibility)."""
    renderer = _owning_renderer(query_func)
    dev = next(renderer.parameters()).device if renderer is not None else torch.device("cuda")
    res = int(resolution)
    lin = torch.stack([torch.linspace(float(bound_min[i]), float(bound_max[i]), res) for i in range(3)]).to(dev)
    lin = lin.contiguous()
    total = res ** 3
    u = torch.empty(total, dtype=torch.float32, device=dev)
    if renderer is not None:
        from . import engine as eng
        w = renderer._prepare()
    pts = torch.empty(min(chunk_points, total), 3, dtype=torch.float32, device=dev)
    for start in range(0, total, chunk_points):
        n = min(chunk_points, total - start)
        call("nunerf_grid_points", res, start, n, lin.data_ptr(), pts.data_ptr())
        if renderer is not None:
            sdf = eng.sdf_infer(w.sdf, pts[:n], w.planes)
        else:
            sdf = query_func(pts[:n]).reshape(-1).float().contiguous()
        call("nunerf_grid_mask", pts.data_ptr(), sdf.data_ptr(), sdf.stride(0), n, float(outside_val),
             u.data_ptr() + 4 * start)
    u = u.view(res, res, res)
    return u if return_device else u.cpu().numpy()


_MC_TABLES = {}


def _mc_tables(dev):
    if dev not in _MC_TABLES:
        from .mc_tables import build_tables
        _MC_TABLES[dev] = tuple(torch.from_numpy(t).to(dev).contiguous() for t in build_tables())
    return _MC_TABLES[dev]


@torch.no_grad()
def marching_cubes(u, isovalue):
    """Device replacement of `mcubes.marching_cubes(u, isovalue)` (field.py:1312): u[res,res,res] float32 (a CUDA tensor
    or a numpy array) -> (vertices float64 [V,3] in grid-index coordinates, triangles int64 [T,3]), numpy, as PyMCubes
    returns them.  Count pass -> scan of the block totals -> emit pass (csrc/mcubes.cu); vertices that sit on the same grid
    edge are merged by their edge id.  Triangles are oriented like PyMCubes': normals point towards DECREASING u (which
    is why extract_mesh_stage1.py:41 flips the faces of an SDF grid).  PyMCubes is neither pinned by the reference nor
    present here: the case tables are generated (mc_tables.py), so the triangulation of ambiguous cells may differ from
    PyMCubes' while the surface is the same (parity unpinned; the tests check the surface properties instead)."""
    if not torch.is_tensor(u):
        u = torch.from_numpy(np.ascontiguousarray(u, dtype=np.float32)).cuda()
    u = u.float().contiguous()
    res = u.shape[0]
    if u.dim() != 3 or u.shape[1] != res or u.shape[2] != res:
        raise ValueError("marching_cubes: u must be a cube [res, res, res]")
    if res < 2:
        return np.zeros((0, 3)), np.zeros((0, 3), dtype=np.int64)
    dev = u.device
    tri_table, n_tris, edges, edge_axis = _mc_tables(dev)
    n_blocks = int(_lib.lib.nunerf_mc_blocks(res))
    counts = torch.empty(n_blocks, dtype=torch.int32, device=dev)
    call("nunerf_mc_count", u.data_ptr(), res, float(isovalue), n_tris.data_ptr(), counts.data_ptr())
    incl = torch.cumsum(counts, 0, dtype=torch.int64)
    offsets = (incl - counts).contiguous()
    total = int(incl[-1])
    if total == 0:
        return np.zeros((0, 3)), np.zeros((0, 3), dtype=np.int64)
    verts = torch.empty(3 * total, 3, dtype=torch.float32, device=dev)
    keys = torch.empty(3 * total, dtype=torch.int64, device=dev)
    call("nunerf_mc_emit", u.data_ptr(), res, float(isovalue), tri_table.data_ptr(), tri_table.shape[1] // 3,
         n_tris.data_ptr(), edges.data_ptr(), edge_axis.data_ptr(), offsets.data_ptr(), verts.data_ptr(), keys.data_ptr())
    uniq, inverse = torch.unique(keys, return_inverse=True)
    vertices = torch.empty(uniq.numel(), 3, dtype=torch.float32, device=dev)
    vertices[inverse] = verts                       # equal keys carry bit-identical positions
    triangles = inverse.view(-1, 3)[:, [0, 2, 1]]   # tables wind towards larger u; PyMCubes towards smaller
    return vertices.double().cpu().numpy(), triangles.contiguous().cpu().numpy()


def extract_geometry(bound_min, bound_max, resolution, threshold, query_func, outside_val=1.0):
    """field.py:1310-1319 with the grid and the marching cubes both on the device (one device->host copy of the
    mesh instead of the res^3 field)."""
    u = extract_fields(bound_min, bound_max, resolution, query_func, outside_val=outside_val, return_device=True)
    vertices, triangles = marching_cubes(u, threshold)
    b_max, b_min = np.asarray(bound_max, dtype=np.float64), np.asarray(bound_min, dtype=np.float64)
    vertices = vertices / (resolution - 1.0) * (b_max - b_min)[None, :] + b_min[None, :]
    return vertices, triangles

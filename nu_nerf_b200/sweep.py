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


def extract_geometry(bound_min, bound_max, resolution, threshold, query_func, outside_val=1.0):
    """field.py:1310-1319.  Marching cubes itself (PyMCubes) is outside the hot path."""
    import mcubes
    u = extract_fields(bound_min, bound_max, resolution, query_func, outside_val=outside_val)
    vertices, triangles = mcubes.marching_cubes(u, threshold)
    b_max, b_min = np.asarray(bound_max, dtype=np.float64), np.asarray(bound_min, dtype=np.float64)
    vertices = vertices / (resolution - 1.0) * (b_max - b_min)[None, :] + b_min[None, :]
    return vertices, triangles

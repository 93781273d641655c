"""Config 5 of BASELINE.json: stage-1 EVAL render (is_train=False, 4096-ray chunks: ray_rgb, depth, normal -- ZT:614-655)
and the extract_fields N^3 SDF sweep (field.py:1286-1307) on the fused inference kernel.  Prints device times (CUDA
events), the sweep's end-to-end time including the single D2H copy, and the tensor-pipe fraction."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nu_nerf_b200.renderer_zerothick import NeROShapeRenderer, load_default_cfg  # noqa: E402
from nu_nerf_b200.sweep import extract_fields, marching_cubes  # noqa: E402
from nu_nerf_b200._lib import call  # noqa: E402

M_SDF_HEAD = 459008


def run(res=512):
    cfg = load_default_cfg()
    cfg["precision"] = "bf16"
    torch.manual_seed(0)
    net = NeROShapeRenderer(cfg, training=False).cuda()
    bmin, bmax = -torch.ones(3), torch.ones(3)
    extract_fields(bmin, bmax, 64, net.sdf_network.sdf)          # warm-up
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    u = extract_fields(bmin, bmax, res, net.sdf_network.sdf, return_device=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    t0 = time.perf_counter()
    uh = extract_fields(bmin, bmax, res, net.sdf_network.sdf)
    wall = time.perf_counter() - t0
    # marching cubes on the resident grid: the two kernels alone (HBM: the grid is read once per pass) and the whole call
    from nu_nerf_b200 import sweep as sw
    tri_table, n_tris, edges, edge_axis = sw._mc_tables(u.device)
    from nu_nerf_b200 import _lib
    counts = torch.empty(int(_lib.lib.nunerf_mc_blocks(res)), dtype=torch.int32, device=u.device)
    call("nunerf_mc_count", u.data_ptr(), res, 0.0, n_tris.data_ptr(), counts.data_ptr())
    torch.cuda.synchronize()
    e0.record()
    call("nunerf_mc_count", u.data_ptr(), res, 0.0, n_tris.data_ptr(), counts.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    ms_count = e0.elapsed_time(e1)
    marching_cubes(u, 0.0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mv, mt = marching_cubes(u, 0.0)
    mc_wall = time.perf_counter() - t0
    mc = {"count_pass_ms": ms_count, "count_pass_GBps": 4.0 * res ** 3 / ms_count / 1e6,
          "call_wall_s_incl_vertex_merge_and_d2h": mc_wall, "vertices": int(len(mv)), "triangles": int(len(mt))}
    # eval render: sphere-bounded near/far, no perturbation, validation outputs (depth, normal)
    from nu_nerf_b200 import synthetic as orc
    R = 4096
    o, d = (t.cuda() for t in orc.synthetic_rays(R))
    near, far = net.near_far_from_sphere(o, d)

    def eval_render():
        with torch.no_grad():
            return net.render(o, d, near, far, None, 0, 0.2, is_train=False, step=10000, is_nerf=False)
    for _ in range(3):
        out = eval_render()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        out = eval_render()
    e1.record()
    torch.cuda.synchronize()
    ms_eval = e0.elapsed_time(e1) / 10
    assert out["depth"].shape == (R, 1) and out["normal"].shape == (R, 3)
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["bf16_tflops_sustained"] \
        if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 1400.0
    tf = 2.0 * M_SDF_HEAD * res ** 3 / (ms * 1e-3) / 1e12
    return {"workload": f"extract_fields {res}^3 SDF sweep (bf16 fused chain)", "device_ms": ms,
                      "end_to_end_s_with_d2h": wall, "points": res ** 3, "Mpts_per_s": res ** 3 / ms / 1e3,
                      "tflops": tf, "frac_of_tensor_peak": tf / peak, "inside_fraction": float((uh < 1.0).mean()),
                      "marching_cubes": mc,
                      "eval_render": {"rays": R, "ms": ms_eval, "rays_per_s": R / ms_eval * 1e3,
                                      "outputs": "ray_rgb, depth, normal, acc, color_bkgr, color_spec (is_train=False)"}}


if __name__ == "__main__":
    print(json.dumps(run(int(sys.argv[1]) if len(sys.argv) > 1 else 512)))

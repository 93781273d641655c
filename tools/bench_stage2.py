"""Config 4 of BASELINE.json: stage-2 nested refraction on a ~100k-triangle synthetic outer mesh (UV sphere r 0.6,
224 x 224 -> 99 904 triangles) + random-init nested inner field.  Reports trace-only Mrays/s (BVH closest hit +
re-intersection, 32 B/ray/bounce algorithmic), the full forward rays/s (ray_trace + render_core, bf16 mode) and the
stage-2 TRAINING step rays/s (ray_trace + render_core forward + backward w.r.t. the field parameters + Adam)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from nu_nerf_b200.synthetic import make_stage2, uv_sphere  # noqa: E402


def timeit(fn, iters=10, warm=3, per_iter=False):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if per_iter:
        # one event pair per iteration: the training step draws new sample positions every iteration, its list sizes
        # change, and an iteration that makes the caching allocator call cudaMalloc costs several times a normal one
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        for a, b in ev:
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        return sorted(a.elapsed_time(b) for a, b in ev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def run(R=4096, train_iters=20, thick=False):
    """thick: the non-zero-thickness renderer of network/renderer.py (nu_nerf_b200/renderer.py) on the same scene."""
    V, Fc = uv_sphere(0.6, 224, 224)
    net = make_stage2("bf16", mesh=(V, Fc), thick=thick).cuda()
    # render(rays_o, rays_d, [mask,] near, far, human_poses, ...): the NZ signature carries the extra mask argument
    rend = (lambda **kw: net.render(o, d, None, None, None, None, -1, 0.2, **kw)) if thick else \
        (lambda **kw: net.render(o, d, None, None, None, -1, 0.2, **kw))
    g = torch.Generator().manual_seed(1)
    o = 3.0 * torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1)
    d = torch.nn.functional.normalize(-o + 0.3 * torch.randn(R, 3, generator=g), dim=-1)
    o, d = o.cuda(), d.cuda()
    net._prepare()
    bvh = net.scene.optix_mesh.bvh
    # trace-only: many rays so that the kernel is not launch bound
    Rt = 1 << 20
    ot = 3.0 * torch.nn.functional.normalize(torch.randn(Rt, 3, device="cuda"), dim=-1)
    dt = torch.nn.functional.normalize(-ot + 0.3 * torch.randn(Rt, 3, device="cuda"), dim=-1)
    ms_trace = timeit(lambda: bvh.trace(ot, dt))
    ms_di = timeit(lambda: net.scene.Dintersect(ot, dt))
    hit, _ = bvh.trace(ot, dt)
    with torch.no_grad():
        ms_full = timeit(lambda: rend(is_train=False, step=10000, is_nerf=True),
                         iters=5, warm=2)
        ms_rt = timeit(lambda: net.ray_trace(o, d), iters=5, warm=2)
    # training step: trainer loss of configs/stage2/nerf/spherepot.yaml (TIR-masked charbonnier + 0.02 eikonal), Adam
    gt = torch.rand(R, 3, generator=g).cuda()
    opt = torch.optim.Adam([p for p in net.parameters() if p.requires_grad], lr=5e-4, fused=True)

    def train_step():
        opt.zero_grad(set_to_none=True)
        out = rend(is_train=True, step=10000, is_nerf=True)
        tm = out["tir_mask"].detach()
        loss = net.compute_rgb_loss(out["ray_rgb"] * tm, gt * tm).mean() + (0.02 * out["gradient_error"]).mean()
        loss.backward()
        opt.step()
        return loss
    t_train = timeit(train_step, iters=train_iters, warm=6, per_iter=True)
    ms_train, ms_train_mean = t_train[len(t_train) // 2], sum(t_train) / len(t_train)
    net.cfg["frozen_ior"] = True            # geometry constant: no replay, no position-gradient kernels
    t_frozen = timeit(train_step, iters=train_iters, warm=3, per_iter=True)
    net.cfg["frozen_ior"] = False
    ms_frozen = t_frozen[len(t_frozen) // 2]
    return {
        "workload": f"stage-2 {'non-zero-thickness (network/renderer.py)' if thick else 'zero-thickness'} forward, outer mesh {Fc.shape[0]} triangles ({bvh.n_nodes} BVH4 nodes), "
                    f"{R} rays, bf16 mode",
        "trace_only": {"rays": Rt, "ms": ms_trace, "Mrays_per_s": Rt / ms_trace / 1e3,
                       "algorithmic_GBs": 32.0 * Rt / ms_trace / 1e6, "hit_fraction": float(hit.mean())},
        "trace_plus_reintersection": {"ms": ms_di, "Mrays_per_s": Rt / ms_di / 1e3},
        "ray_trace_with_sampling": {"rays": R, "ms": ms_rt, "rays_per_s": R / ms_rt * 1e3},
        "full_forward": {"rays": R, "ms": ms_full, "rays_per_s": R / ms_full * 1e3},
        "train_step": {"rays": R, "ms": ms_train, "rays_per_s": R / ms_train * 1e3, "ms_mean": ms_train_mean,
                       "ms_min_max": [t_train[0], t_train[-1]], "timing": f"median of {train_iters} per-step CUDA-event timings",
                       "note": "forward + backward w.r.t. all parameters (incl. IORs_pred through the path geometry) + Adam"},
        "train_step_frozen_ior": {"rays": R, "ms": ms_frozen, "rays_per_s": R / ms_frozen * 1e3,
                                  "note": "cfg['frozen_ior']: path geometry constant, IORs_pred not trained"},
    }


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    print(json.dumps(run(int(args[0]) if args else 4096, thick="--thick" in sys.argv)))

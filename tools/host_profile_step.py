"""Host-side cost of one stage-1 training step (trainer.step): enqueue-only wall time vs device time, and a cProfile
of the Python side.  `python tools/host_profile_step.py [step]` (default 20000: occlusion-probe loss active)."""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from nu_nerf_b200 import dist as nd, synthetic as syn  # noqa: E402
from nu_nerf_b200.renderer_zerothick import NeROShapeRenderer, load_default_cfg  # noqa: E402


def main():
    step = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    R = 4096
    dev = torch.device("cuda", 0)
    cfg = load_default_cfg()
    cfg["precision"] = "bf16"
    cfg["train_ray_num"] = R
    torch.manual_seed(0)
    net = NeROShapeRenderer(cfg, training=False).to(dev)
    anneal = float(net.get_anneal_val(step))
    render_fn = lambda o, d, n, f, s: net.render(o, d, n, f, None, -1, anneal, is_train=True, step=s, is_nerf=True)
    trainer = nd.DataParallelTrainer(net, render_fn, net.compute_rgb_loss, lr_fn=lambda s: bench.lr_at(s),
                                     eikonal_weight=bench.EIK_W, occ_loss_step=cfg["occ_loss_step"])
    o, d = (t.to(dev) for t in syn.synthetic_rays(R))
    gt = syn.synthetic_targets(R).to(dev)
    near, far = torch.full((R, 1), 0.8, device=dev), torch.full((R, 1), 4.5, device=dev)
    run = lambda: trainer.step(o, d, gt, near, far, step, chunk=R)
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    n = 10
    t0 = time.perf_counter()
    for _ in range(n):
        run()
    t_enq = (time.perf_counter() - t0) / n
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t0) / n
    t0 = time.perf_counter()
    for _ in range(n):
        run().item()
    t_sync = (time.perf_counter() - t0) / n
    print(f"step {step}: host enqueue {t_enq * 1e3:.2f} ms/step, pipelined {t_all * 1e3:.2f} ms/step, "
          f"with a loss read-back every step {t_sync * 1e3:.2f} ms/step")
    # every synchronising call of one step, with its Python location
    import warnings
    torch.cuda.set_sync_debug_mode("warn")
    with warnings.catch_warnings(record=True) as ws:
        warnings.simplefilter("always")
        run()
    torch.cuda.set_sync_debug_mode("default")
    for wn in ws:
        print(f"SYNC at {wn.filename.split('/repo/')[-1]}:{wn.lineno}: {str(wn.message)[:90]}")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(5):
        run().item()
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(22)
    # which aten ops / CUDA runtime calls hold the host (synchronising copies show up as cudaMemcpyAsync / cudaStreamSynchronize)
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            run().item()
    rows = sorted(prof.key_averages(), key=lambda e: -e.self_cpu_time_total)[:18]
    for e in rows:
        print(f"{e.self_cpu_time_total / 3:9.0f} us/step self-cpu {e.count // 3:5d}x  {e.key[:80]}")


if __name__ == "__main__":
    main()

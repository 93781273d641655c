#!/usr/bin/env python
"""bench.py -- stage-1 training-step throughput (rays/s) of the NU-NeRF hot path on B200.

  python bench.py --gpus N --steps K --warmup W            our arm (one process per GPU; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  the UNMODIFIED reference on the host cores (oracle/_ref staged by
                                                           oracle/make_ref.py; the oracle port only if that tree is absent)

A step = sample_ray + render_core forward + trainer loss + backward + (NCCL grad all-reduce) + Adam on one batch of
synthetic rays (spherepot field, random init, step 10000, rays on a radius-3 sphere).  `value` is measured
with the ray batch already resident in HBM; `e2e` repeats the measurement through the renderer's public API with the
batch in pinned host memory (H2D of rays/targets and a D2H read of the loss inside the timed region).

At N = 1 the line also carries `configs`: the other BASELINE.json configurations measured inside the same command --
the fp32-accurate split mode on config 2, config 3's 32 768-ray batch on one GPU, config 4 (stage-2 trace + training
step on the 99 904-triangle mesh) and config 5 (eval render, 512^3 sweep, marching cubes).  At N > 1 the default batch is
config 3's 32 768 rays per GPU and the 4096-ray measurement is kept as `configs.config2_rays4096`.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STEP = 10000                      # cos_anneal 0.2, inv_s frozen, no occ loss
EIK_W = 0.1
M_SDF, M_SDF_HEAD, M_NERF, M_COL, M_OL = 524544, 459008, 604160, 1865472, 150272   # MACs / point


def lr_at(step, lr=5e-4, warm=5000, end=300000, alpha=0.05):
    """train/lr_common_manager.py:36-46 (WarmUpCosLR)."""
    if step < warm:
        return lr * step / warm
    prog = (step - warm) / (end - warm)
    return lr * ((math.cos(math.pi * prog) + 1.0) * 0.5 * (1 - alpha) + alpha)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# =============================================================================================== reference arm
def _reference_root():
    """The unmodified reference's files: /root/reference in the build container, oracle/_ref (staged by
    oracle/make_ref.py at build time, git-ignored, shipped with the working tree) on the GPU box."""
    for r in (os.environ.get("NUNERF_REFERENCE_ROOT"), os.path.join(ROOT, "oracle", "_ref"), "/root/reference"):
        if r and os.path.isdir(os.path.join(r, "network")):
            return r
    return None


def cpu_step_fn(R):
    """One training step of the reference's CPU implementation on R rays (all host threads); returns (callable, kind).
    kind "reference": the UNMODIFIED reference module (network/renderer_zerothick.py NeROShapeRenderer.render + the
    trainer's loss + torch Adam) imported through the shim layer of oracle/ref_harness.py; kind "port": the oracle
    restatement, only when no reference tree is available."""
    import torch
    torch.set_num_threads(os.cpu_count())
    root = _reference_root()
    if root is not None:
        os.environ["NUNERF_REFERENCE_ROOT"] = root
        from oracle import ref_harness as rh
        rh.REF_ROOT = root
        net, _ = rh.load_stage1(seed=0)
        o, d = rh.synthetic_rays(R)
        gt = rh.synthetic_targets(R)
        near, far = torch.full((R, 1), 0.8), torch.full((R, 1), 4.5)
        poses = torch.eye(3, 4)[None].repeat(R, 1, 1)
        opt = torch.optim.Adam(net.parameters(), lr=lr_at(STEP))

        def step():
            opt.zero_grad(set_to_none=True)
            with rh.in_ref_dir():
                out = net.render(o, d, near, far, poses, -1, net.get_anneal_val(STEP), is_train=True, step=STEP, is_nerf=True)
            loss = net.compute_rgb_loss(out["ray_rgb"], gt).mean() + (EIK_W * out["gradient_error"]).mean()
            if STEP >= net.cfg["occ_loss_step"]:
                loss = loss + out["loss_occ"].mean() + 0.5 * torch.nn.functional.mse_loss(out["color_bkgr"].flatten(),
                                                                                          out["color_spec"].flatten())
            loss.backward()
            opt.step()
            return float(loss)
        return step, "reference"
    from oracle import nunerf_oracle as orc
    from nu_nerf_b200.renderer_zerothick import NeROShapeRenderer, load_default_cfg
    torch.manual_seed(0)
    net = NeROShapeRenderer(load_default_cfg(), training=False)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    params = {k: v.requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point and k != "color_network.FG_LUT"}
    opt = torch.optim.Adam(list(params.values()), lr=lr_at(STEP))
    o, d = orc.synthetic_rays(R)
    gt = orc.synthetic_targets(R)
    near, far = torch.full((R, 1), 0.8), torch.full((R, 1), 4.5)

    def step():
        U0, U1 = torch.rand(R, 1), torch.rand(R, 32)
        opt.zero_grad(set_to_none=True)
        out = orc.render(sd, o, d, near, far, U0, U1, 0.2, STEP)
        loss = orc.train_loss(out, gt, EIK_W, step=STEP)
        loss.backward()
        opt.step()
        return float(loss)
    return step, "port"


def _cpu_what(kind):
    return ("unmodified reference (network/renderer_zerothick.py through oracle/ref_harness.py shims) on torch CPU"
            if kind == "reference" else "oracle port (oracle/nunerf_oracle.py) on torch CPU")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    R = args.cpu_rays
    step, kind = cpu_step_fn(R)
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = R / dt
    cores = os.cpu_count()
    line = {
        "impl": "reference", "metric": "train-step rays/sec", "value": val, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"stage-1 spherepot training step (sample_ray + render_core fwd+bwd + Adam), step {STEP}, "
                               f"bounded sample of {R} rays per step, {_cpu_what(kind)}"},
        "cpu_baseline": {"value": val, "unit": "rays/s", "cores": cores, "kind": kind,
                         "sample": f"{R} rays x {args.steps} steps, {cores} threads, {_cpu_what(kind)}"},
        "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# =============================================================================================== our arm
def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of `kernel` from the committed `ncu --set full`
    capture (profiles/ncu_traffic.json: written by tools/ncu_traffic.py from the .ncu-rep, names the capture)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        t = json.load(f).get(kernel)
    return t if t else None


def tc_work(name, a):
    """(kernel family, key, algorithmic FLOPs, algorithmic HBM bytes) of one tensor-core entry-point call."""
    if name == "nunerf_linear":
        p = a[0]._obj
        n = p.n_store if p.n_store else p.N
        key = (f"linear K={p.K} N={p.N} act={p.act} aux={p.aux_mode} add={int(bool(p.add))} "
               f"mask_out={int(bool(p.mask_out))} f32={int(bool(p.out_f32))} bf16={int(bool(p.out))} lo={int(bool(p.a_lo_off))}")
        return "linear_tc_kernel", key, 2.0 * p.M * n * p.K, 2.0 * p.M * (p.K + (n if p.out else 0)) * (2 if p.a_lo_off else 1)
    if name == "nunerf_linear_dw":
        p = a[0]._obj
        return "dw_tc_kernel", f"dw K={p.K} N={p.N}", 2.0 * p.M * p.N * p.K, 2.0 * p.M * (p.K + p.N) * (2 if p.z_lo_off else 1)
    if name == "nunerf_mlp_chain":
        p = a[0]._obj
        fl, by, desc = 0.0, 2.0 * p.M * p.K0, []
        for l in range(p.n_layers):
            L = p.layer[l]
            fl += 2.0 * p.M * L.N * L.K
            # stores, masks, fp32 heads and -- SDF reverse-pass layers -- the stored activation / adjoint they read back
            # (aux1, aux2) and the E term they write: all of it is traffic the algorithm requires of the launch
            by += p.M * ((2.0 * L.N if L.store else 0.0) + (L.N / 8.0 if L.mask_out else 0.0)
                         + (L.N / 8.0 if L.mask_in else 0.0) + (4.0 * L.n32 if L.out32 else 0.0)
                         + 2.0 * L.N * (int(bool(L.aux1)) + int(bool(L.aux2)) + int(bool(L.e_out))))
            desc.append(f"{L.K}>{L.N}" + ("" if not L.aux_mode else f"[aux{L.aux_mode}]"))
        cls = "sdf_reverse_chains" if any(p.layer[l].aux_mode >= 4 for l in range(p.n_layers)) else "plain_chains"
        return "mlp_chain_kernel", "chain " + " ".join(desc), fl, by, cls
    if name == "nunerf_sdf_infer":
        p = a[0]._obj
        return "mlp_chain_kernel", "chain sdf_infer (PE + 9 layers)", 2.0 * p.M * M_SDF_HEAD, 16.0 * p.M, "sdf_infer"
    return None


def hbm_work(name, a):
    """Algorithmic HBM bytes of one launch of the sampling / compositing kernels (fp32 I/O that the kernel must do:
    its inputs once, its outputs once; SURVEY 8d conventions)."""
    if name == "nunerf_composite_fwd":
        R, S = a[5], a[6]
        return 16.0 * R * S + 28.0 * R
    if name == "nunerf_composite_bwd":
        R, S = a[5], a[6]
        return 32.0 * R * S + 12.0 * R
    if name == "nunerf_upsample":
        R, n, nn = a[4], a[5], a[6]
        return R * (24.0 + 8.0 * n + 8.0 * nn + 8.0 * (n + nn))
    if name == "nunerf_points":
        R, n = a[3], a[4]
        return R * (24.0 + 16.0 * n)
    if name == "nunerf_merge_sdf":
        R, n, nn = a[3], a[4], a[5]
        return R * 4.0 * (n + nn + 2 * (n + nn))
    if name == "nunerf_ray_setup":
        R = a[7]
        return R * (24.0 + 8.0 + 132.0 + 4.0 * 96)
    if name == "nunerf_render_geometry":
        R, S = a[3], a[4]
        # o, d + z in; compact points / dists / dirs (28 B per sample) + the per-ray map (40 B) out
        return R * (24.0 + 4.0 * S + 28.0 * S + 40.0)
    return 0.0


def profile_step(train_step, inputs, ops, torch):
    """Per-entry-point device time of one step (CUDA events around every C-ABI call; printed to stderr)."""
    import collections
    rec = []
    orig_call = ops.call
    from nu_nerf_b200 import engine as eng

    def timed_call(name, *a):
        w = tc_work(name, a)
        if w is not None:
            key, nbytes = w[1], w[3]
        else:
            key, nbytes = name, hbm_work(name, a)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig_call(name, *a)
        e1.record()
        rec.append((key, e0, e1, nbytes))
    ops.call = timed_call
    eng.call = timed_call
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    try:
        t0.record()
        train_step(*inputs)
        t1.record()
        torch.cuda.synchronize()
    finally:
        ops.call = orig_call
        eng.call = orig_call
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for key, e0, e1, nb in rec:
        a = agg[key]
        a[0] += 1; a[1] += e0.elapsed_time(e1); a[2] += nb
    tot = sum(v[1] for v in agg.values())
    print(f"[profile] step {t0.elapsed_time(t1):.2f} ms, inside C-ABI calls {tot:.2f} ms", file=sys.stderr)
    for key, (n, ms, nb) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        gbs = f"{nb / ms / 1e6:8.0f} GB/s" if nb else ""
        print(f"[profile] {ms:8.3f} ms {n:4d}x  {key} {gbs}", file=sys.stderr)


def two_phase_fns(net, anneal):
    """Callbacks that let the trainer know the step's global inner-sample count before differentiating a chunk (exact
    eikonal mean for chunked / sharded steps, nu_nerf_b200/dist.py)."""
    return {"sample_fn": lambda o, d, n_, f_, st: net.sample_ray(o, d, n_, f_, net.cfg["perturb"]),
            "core_fn": lambda o, d, z, st: net.render_core(o, d, z, None, cos_anneal_ratio=anneal, step=st, is_train=True,
                                                           is_nerf=True),
            "count_fn": lambda o, d, z: net.count_inner(o, d, z)}


def stage1_quick(args, torch, dist, dev, precision, R, steps, world=1, rank=0):
    """A short measurement of the stage-1 training step at another precision / batch size (sub-records of the line)."""
    from nu_nerf_b200 import dist as nd
    from nu_nerf_b200 import synthetic as syn
    from nu_nerf_b200.renderer_zerothick import NeROShapeRenderer, load_default_cfg
    cfg = load_default_cfg()
    cfg["precision"] = precision
    cfg["train_ray_num"] = R
    torch.manual_seed(0)
    net = NeROShapeRenderer(cfg, training=False).to(dev)
    anneal = float(net.get_anneal_val(STEP))
    trainer = nd.DataParallelTrainer(net, lambda o, d, n_, f_, st: net.render(o, d, n_, f_, None, -1, anneal, is_train=True,
                                                                             step=st, is_nerf=True),
                                     net.compute_rgb_loss, lr_fn=lambda s_: lr_at(s_), eikonal_weight=EIK_W,
                                     occ_loss_step=cfg["occ_loss_step"], **two_phase_fns(net, anneal))
    o_all, d_all = syn.synthetic_rays(R * world, seed=1)
    gt_all = syn.synthetic_targets(R * world, seed=3)
    sel = nd.shard_batch(torch.arange(R * world), rank, world)
    o, d, gt = (t[sel].contiguous().to(dev) for t in (o_all, d_all, gt_all))
    near, far = torch.full((R, 1), 0.8, device=dev), torch.full((R, 1), 4.5, device=dev)
    chunk = min(args.chunk, R)
    for _ in range(3):
        trainer.step(o, d, gt, near, far, STEP, chunk=chunk)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        trainer.step(o, d, gt, near, far, STEP, chunk=chunk)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = ms.item() / steps
    n_in = trainer.last["n_in"]
    flops = (R / chunk) * (2.0 * chunk * 112 * M_SDF_HEAD + n_in * 2.0 * (3 * M_SDF + 3 * M_SDF_HEAD + 3 * M_COL)
                           + (chunk * 160 - n_in) * 2.0 * 3 * M_NERF + chunk * 2.0 * 3 * M_OL)
    del trainer, net
    torch.cuda.empty_cache()
    return {"precision": precision, "rays_per_gpu": R, "n_gpus": world, "steps": steps, "ms_per_step": ms,
            "value": R * world / (ms * 1e-3), "unit": "rays/s", "step_algorithmic_tflops_per_gpu": flops / (ms * 1e-3) / 1e12}


def sub_records(args, torch, dev, hbm, tf_sus):
    """The other BASELINE.json configurations, measured inside the same command (N = 1)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    out = {}
    try:
        sp = stage1_quick(args, torch, None, dev, "split", args.rays_per_gpu, steps=3)
        sp["frac_of_tensor_peak"] = sp["step_algorithmic_tflops_per_gpu"] / tf_sus
        sp["note"] = ("fp32-accurate mode (bf16 hi + lo planes, 3 MMAs per product, layer-by-layer kernels): the mode "
                      "that meets the 1e-4 rgb / 1e-3 gradient gates")
        out["split"] = sp
        c3 = stage1_quick(args, torch, None, dev, "bf16", 32768, steps=3)
        c3["frac_of_tensor_peak"] = c3["step_algorithmic_tflops_per_gpu"] / tf_sus
        out["config3_rays32768_1gpu"] = c3
        import bench_stage2
        c4 = bench_stage2.run(4096, train_iters=12)
        c4["trace_only"]["hbm_frac_algorithmic"] = c4["trace_only"]["algorithmic_GBs"] / hbm
        out["config4_stage2"] = c4
        torch.cuda.empty_cache()
        # the same scene through the non-zero-thickness renderer (network/renderer.py)
        nz = bench_stage2.run(4096, train_iters=8, thick=True)
        out["f1_stage2_nonzero_thickness"] = {k: nz[k] for k in ("workload", "ray_trace_with_sampling", "full_forward",
                                                                 "train_step", "train_step_frozen_ior")}
        torch.cuda.empty_cache()
        import bench_sweep
        out["config5_sweep_eval"] = bench_sweep.run(512)
        torch.cuda.empty_cache()
    except Exception as e:            # a sub-record must never take the headline down with it
        out["error"] = f"{type(e).__name__}: {e}"
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from nu_nerf_b200 import _lib, ops
    from nu_nerf_b200 import dist as nd
    from nu_nerf_b200.renderer_zerothick import NeROShapeRenderer, load_default_cfg
    from nu_nerf_b200 import synthetic as orc  # seeded synthetic rays / targets (the product arm never touches oracle/)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.rays_per_gpu <= 0:          # default: config 2 on one GPU, config 3 (32 768 rays per GPU) on several
        args.rays_per_gpu = 4096 if world == 1 else 32768
    R = args.rays_per_gpu
    chunk = min(args.chunk, R)
    cfg = load_default_cfg()
    cfg["precision"] = args.precision
    cfg["train_ray_num"] = R
    torch.manual_seed(0)
    net = NeROShapeRenderer(cfg, training=False).to(dev)
    anneal = float(net.get_anneal_val(STEP))

    def render_fn(o, d, near_, far_, step):
        return net.render(o, d, near_, far_, None, -1, anneal, is_train=True, step=step, is_nerf=True)
    # the product's ray-sharded trainer: global-denominator losses, one all-reduce of the flat gradient, CUDA Adam
    trainer = nd.DataParallelTrainer(net, render_fn, net.compute_rgb_loss, lr_fn=lambda s: lr_at(s), eikonal_weight=EIK_W,
                                     occ_loss_step=cfg["occ_loss_step"], **two_phase_fns(net, anneal))
    flat = trainer.fp.flat
    # rays: the same generator on every rank, rank-strided slices of one global batch
    o_all, d_all = orc.synthetic_rays(R * world, seed=1)
    gt_all = orc.synthetic_targets(R * world, seed=3)
    sel = nd.shard_batch(torch.arange(R * world), rank, world)
    o_h, d_h, gt_h = (t[sel].contiguous().pin_memory() for t in (o_all, d_all, gt_all))
    o_d, d_d, gt_d = o_h.to(dev), d_h.to(dev), gt_h.to(dev)
    near = torch.full((R, 1), 0.8, device=dev)
    far = torch.full((R, 1), 4.5, device=dev)
    stats = {"n_in": 0, "n_out": 0}

    def train_step(o, d, gt):
        total = trainer.step(o, d, gt, near, far, STEP, chunk=chunk)
        stats["n_in"] = trainer.last["n_in"]
        return total

    def e2e_step():
        o = o_h.to(dev, non_blocking=True)
        d = d_h.to(dev, non_blocking=True)
        gt = gt_h.to(dev, non_blocking=True)
        return float(train_step(o, d, gt).item())      # D2H read of the step's loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / n

    for _ in range(max(args.warmup, 3)):
        train_step(o_d, d_d, gt_d)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(1.5)          # nvidia-smi start-up takes driver locks; keep it out of the timed region
    for _ in range(2):
        train_step(o_d, d_d, gt_d)
    l0 = _lib.launch_count()
    ms = timed(lambda: train_step(o_d, d_d, gt_d), args.steps)
    launches = (_lib.launch_count() - l0) // args.steps
    for _ in range(max(args.warmup, 3)):       # the end-to-end path gets its own untimed warm-up (first pinned H2D copies and
        e2e_step()                             # the first .item() of a process pay one-off driver initialisation)
    ms_e2e = timed(e2e_step, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # ---- roofline of the tensor-core kernels (linear_tc_kernel, dw_tc_kernel, mlp_chain_kernel), measured with CUDA
    #      events around every launch of one extra step on the launching stream; the dominant one is reported
    rec = []
    orig_call = ops.call
    from nu_nerf_b200 import engine as eng

    def timed_call(name, *a):
        w = tc_work(name, a)
        if w is None:
            return orig_call(name, *a)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig_call(name, *a)
        e1.record()
        rec.append((w[0], e0, e1, w[2], w[3], w[4] if len(w) > 4 else None))
    ops.call = timed_call
    eng.call = timed_call
    try:
        train_step(o_d, d_d, gt_d)
        torch.cuda.synchronize()
    finally:
        ops.call = orig_call
        eng.call = orig_call
    if args.profile:
        profile_step(train_step, (o_d, d_d, gt_d), ops, torch)
    fam, chain_cls = {}, {}
    for k, e0, e1, fl, by, cls in rec:
        dt = e0.elapsed_time(e1)
        f = fam.setdefault(k, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
        f["ms"] += dt; f["flops"] += fl; f["bytes"] += by; f["launches"] += 1
        if cls is not None:
            c = chain_cls.setdefault(cls, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
            c["ms"] += dt; c["flops"] += fl; c["bytes"] += by; c["launches"] += 1
    top = max(fam, key=lambda k: fam[k]["ms"])
    t_lin, fl_lin, by_lin, n_lin = fam[top]["ms"] * 1e-3, fam[top]["flops"], fam[top]["bytes"], fam[top]["launches"]
    hbm, tf_burst, tf_sus, which = peaks()
    n_in, n_out = stats["n_in"], chunk * 160 - stats["n_in"]
    flops_step = (R / chunk) * (2.0 * chunk * 112 * M_SDF_HEAD + n_in * 2.0 * (3 * M_SDF + 3 * M_SDF_HEAD + 3 * M_COL)
                                + n_out * 2.0 * 3 * M_NERF + chunk * 2.0 * 3 * M_OL)
    sub4096 = None
    if world > 1 and R != 4096 and not args.no_subrecords:
        # the 4096-ray batch of config 2 on the same N GPUs, kept beside the config-3 headline (a collective measurement:
        # every rank takes part before the non-zero ranks leave)
        sub4096 = stage1_quick(args, torch, dist, dev, "bf16", 4096, steps=max(3, args.steps // 2), world=world, rank=rank)
    if rank != 0:
        dist.destroy_process_group()
        return
    value = R * world / (ms * 1e-3)
    h2d = int(o_h.numel() + d_h.numel() + gt_h.numel()) * 4
    line = {
        "metric": "train-step rays/sec", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "bf16x3-split(fp32-accurate)",
        "data": "synthetic",
        "config": {"workload": f"stage-1 spherepot training step: sample_ray (64+4x16+32) + render_core fwd+bwd + Adam, "
                               f"{R} rays/GPU (chunks of {chunk}), step {STEP}, random-init field, synthetic rays (SURVEY 8d)",
                   "rays_per_gpu": R, "samples_per_ray": 160, "inner_samples_per_chunk": n_in,
                   "l2": "no flush needed: per-step activation working set (GBs) >> 126 MB L2",
                   "parallelism": f"ray-sharded dp{world}, NCCL all-reduce of {flat.numel()} fp32 grads"},
        "e2e": {"value": R * world / (ms_e2e * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": top + " (tcgen05)",
                     "achieved": fl_lin / t_lin / 1e12 if t_lin > 0 else None, "peak": tf_sus, "unit": "TFLOP/s",
                     "frac": fl_lin / t_lin / 1e12 / tf_sus if t_lin > 0 else None, "traffic": (ncu_traffic(top) or {}).get("dram_bytes_per_launch"),
                     "traffic_note": (ncu_traffic(top) or {}).get("note"),
                     "peak_source": f"bf16_tflops_sustained of {which} (kernel timed inside a long step)",
                     "launches_timed": n_lin, "kernel_share_of_step": t_lin * 1e3 / ms,
                     "kernels": {k: {"ms_per_step": v["ms"], "launches": v["launches"], "share_of_step": v["ms"] / ms,
                                     "tflops": v["flops"] / v["ms"] / 1e9, "frac_of_tensor_peak": v["flops"] / v["ms"] / 1e9 / tf_sus,
                                     "algorithmic_hbm_gbs": v["bytes"] / v["ms"] / 1e6} for k, v in fam.items()},
                     # the chain launches by what bounds them: the fused SDF query (no stores: tensor / epilogue), plain
                     # forward and ReLU-backward chains (one store per layer), the SDF reverse-pass chains (1.5-2 KB per
                     # row and layer in and out: closer to the HBM roofline than to the tensor one)
                     "chain_classes": {k: {"ms_per_step": v["ms"], "launches": v["launches"],
                                           "tflops": v["flops"] / v["ms"] / 1e9,
                                           "frac_of_tensor_peak": v["flops"] / v["ms"] / 1e9 / tf_sus,
                                           "algorithmic_hbm_gbs": v["bytes"] / v["ms"] / 1e6,
                                           "frac_of_hbm_peak": v["bytes"] / v["ms"] / 1e6 / hbm} for k, v in chain_cls.items()},
                     "hbm_achieved_gbs": by_lin / t_lin / 1e9 if t_lin > 0 else None,
                     "hbm_frac": by_lin / t_lin / 1e9 / hbm if t_lin > 0 else None,
                     "step_algorithmic_tflops": flops_step / (ms * 1e-3) / 1e12,
                     "step_frac_of_tensor_peak": flops_step / (ms * 1e-3) / 1e12 / tf_sus},
    }
    if world == 1 and not args.no_subrecords:
        line["configs"] = sub_records(args, torch, dev, hbm, tf_sus)
    if sub4096 is not None:
        line["configs"] = {"config2_rays4096": sub4096}
    if world == 1 and not args.no_cpu_baseline:
        Rc = args.cpu_rays
        step, kind = cpu_step_fn(Rc)
        step()
        t0 = time.perf_counter()
        n = 2
        for _ in range(n):
            step()
        dt = (time.perf_counter() - t0) / n
        line["cpu_baseline"] = {"value": Rc / dt, "unit": "rays/s", "cores": os.cpu_count(), "kind": kind,
                                "sample": f"{Rc} rays x {n} steps (1 warm-up), {os.cpu_count()} threads, {_cpu_what(kind)}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    global STEP
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rays-per-gpu", type=int, default=0, help="default: 4096 (config 2) at N = 1, 32768 (config 3) at N > 1")
    ap.add_argument("--chunk", type=int, default=8192, help="rays per render call (memory bound)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "split"])
    ap.add_argument("--step", type=int, default=STEP, help="training step the schedule is evaluated at (SURVEY 8d: 10000; "
                    "20000 adds the occlusion-probe loss, outer_reg and the trainable inv_s)")
    ap.add_argument("--cpu-rays", type=int, default=256, help="bounded CPU sample (rays per step; config 1 of BASELINE.json "
                    "is 512: per-ray cost is flat in the batch size, SURVEY 6)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-subrecords", action="store_true", help="skip the `configs` sub-records (split mode, configs 3-5)")
    ap.add_argument("--profile", action="store_true", help="print a per-entry-point device-time table to stderr")
    args = ap.parse_args()
    STEP = args.step          # 10000 = the SURVEY 8(d) primary point; 20000 = occlusion-probe loss + outer_reg + trainable inv_s
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

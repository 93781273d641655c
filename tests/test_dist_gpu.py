"""The ray-sharded trainer (nu_nerf_b200/dist.py) with the REAL engine:
  * one GPU: a step split into chunks equals the unsplit step (two-phase global eikonal denominator);
  * two GPUs over NCCL: two ranks with R rays each equal one rank with the 2 R-ray batch -- flat gradient of the first step
    and the weights after two Adam steps (needs >= 2 GPUs: `gpurun --gpus 2`; skipped otherwise).
Deterministic sampling (perturb = 0), fp32-accurate split mode."""
import os
import socket
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu
STEP, R = 10000, 256


def _build(dev):
    from nu_nerf_b200 import dist as nd
    from nu_nerf_b200.renderer_zerothick import NeROShapeRenderer, load_default_cfg
    cfg = load_default_cfg()
    cfg["precision"] = "split"
    cfg["perturb"] = 0.0
    torch.manual_seed(0)
    net = NeROShapeRenderer(cfg, training=False).to(dev)
    anneal = float(net.get_anneal_val(STEP))
    render = lambda o, d, n_, f_, st: net.render(o, d, n_, f_, None, -1, anneal, is_train=True, step=st, is_nerf=True)
    trainer = nd.DataParallelTrainer(
        net, render, net.compute_rgb_loss, eikonal_weight=0.1, occ_loss_step=cfg["occ_loss_step"],
        sample_fn=lambda o, d, n_, f_, st: net.sample_ray(o, d, n_, f_, 0.0),
        core_fn=lambda o, d, z, st: net.render_core(o, d, z, None, cos_anneal_ratio=anneal, step=st, is_train=True, is_nerf=True),
        count_fn=lambda o, d, z: net.count_inner(o, d, z))
    return net, trainer


def _batch(n, dev):
    from nu_nerf_b200 import synthetic as syn
    o, d = syn.synthetic_rays(n, seed=1)
    gt = syn.synthetic_targets(n, seed=3)
    return o.to(dev), d.to(dev), gt.to(dev), torch.full((n, 1), 0.8, device=dev), torch.full((n, 1), 4.5, device=dev)


def _run(trainer, o, d, gt, near, far, chunk, steps=2):
    grads = None
    for s in range(steps):
        trainer.step(o, d, gt, near, far, STEP + s, chunk=chunk)
        if s == 0:
            grads = trainer.fp.grad.clone()
            trainer.n_in_first = trainer.last["n_in_step"]      # (later steps see weights that differ by Adam's sign noise)
    torch.cuda.synchronize()
    return grads.cpu(), trainer.fp.flat.clone().cpu()


def _compare(g_a, w_a, g_b, w_b, lr_steps):
    gs = g_b.abs().max().item()
    assert (g_a - g_b).abs().max().item() < 2e-4 * gs, ((g_a - g_b).abs().max().item(), gs)
    assert (g_a - g_b).norm().item() < 2e-5 * g_b.norm().item()
    dw = (w_a - w_b).abs()
    # Adam divides by sqrt(v): an element whose gradient is ~0 takes a step of either sign -- bounded by lr per step
    assert dw.max().item() <= 2.5 * lr_steps, dw.max().item()
    assert (dw < 1e-5).float().mean().item() > 0.995


def test_chunked_step_equals_the_unsplit_step():
    from nu_nerf_b200 import dist as nd
    dev = torch.device("cuda", 0)
    o, d, gt, near, far = _batch(R, dev)
    _, t1 = _build(dev)
    g1, w1 = _run(t1, o, d, gt, near, far, chunk=R)
    _, t2 = _build(dev)
    g2, w2 = _run(t2, o, d, gt, near, far, chunk=R // 4)
    assert t2.n_in_first == t1.n_in_first > 0
    _compare(g2, w2, g1, w1, 2 * nd.warm_up_cos_lr(STEP))


def _worker(rank, world, port, out_path):
    import torch.distributed as dist
    from nu_nerf_b200 import dist as nd
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        o, d, gt, near, far = _batch(R * world, dev)
        sel = nd.shard_batch(torch.arange(R * world), rank, world).to(dev)
        _, tr = _build(dev)
        g, w = _run(tr, o[sel].contiguous(), d[sel].contiguous(), gt[sel].contiguous(), near[sel], far[sel], chunk=R // 2)
        if rank == 0:
            torch.save({"g": g, "w": w}, out_path)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_nccl_ranks_equal_one_rank_with_the_doubled_batch(tmp_path):
    import torch.multiprocessing as mp
    from nu_nerf_b200 import dist as nd
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "rank0.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    dev = torch.device("cuda", 0)
    o, d, gt, near, far = _batch(2 * R, dev)
    _, tr = _build(dev)
    g, w = _run(tr, o, d, gt, near, far, chunk=2 * R)
    _compare(got["g"], got["w"], g, w, 2 * nd.warm_up_cos_lr(STEP))

"""Host-side logic of the ray-sharded trainer (nu_nerf_b200/dist.py) on CPU: 2 gloo ranks must reproduce the
single-process large-batch step exactly (global denominators, SUM all-reduce of the flat gradient, identical Adam).

The renderer here is a small torch stand-in with the renderer's output contract (`ray_rgb [R,3]`, `gradient_error`
of data-dependent length, `transmission` present iff the rank has inner samples): the CUDA engine itself cannot run
in this container, and the point of this test is the collective / normalisation logic around it.
"""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class TinyField(torch.nn.Module):
    def __init__(self, nz=False):
        super().__init__()
        self.nz = nz            # outputs of the stage-1 class of network/renderer.py: loss_normal, colours on candidate rays
        torch.manual_seed(7)
        self.a = torch.nn.Linear(6, 16)
        self.b = torch.nn.Linear(16, 4)

    def render(self, o, d, near, far, step):
        h = torch.tanh(self.a(torch.cat([o, d], -1)))
        y = self.b(h)
        rgb = torch.sigmoid(y[:, :3])
        # data-dependent number of "inner samples": rays whose 4th output is positive contribute 1..3 samples each
        k = (y[:, 3] > 0).nonzero()[:, 0]
        out = {"ray_rgb": rgb}
        if k.numel() > 0:
            reps = (o[k, 0] > 0).long() + (d[k, 0] > 0).long() + 1       # a property of the ray, not of its index
            g = torch.repeat_interleave(y[k, 3], reps)
            out["gradient_error"] = (g - 1.0) ** 2
            out["transmission"] = g[:, None].detach()
        else:
            out["gradient_error"] = torch.zeros(1)
        if self.nz:
            out["loss_normal"] = torch.relu(y[:, 3:4])
            cand = o[:, 1] > 0                                           # a property of the ray: a data-dependent subset
            out["color_bkgr"], out["color_spec"] = rgb[cand], torch.sigmoid(0.5 * y[cand, :3])
        return out


def torch_adam(fp, lr, betas=(0.9, 0.999), eps=1e-8):
    fp.t += 1
    fp.m.mul_(betas[0]).add_(fp.grad, alpha=1 - betas[0])
    fp.v.mul_(betas[1]).addcmul_(fp.grad, fp.grad, value=1 - betas[1])
    mhat = fp.m / (1 - betas[0] ** fp.t)
    vhat = fp.v / (1 - betas[1] ** fp.t)
    fp.flat.addcdiv_(mhat, vhat.sqrt() + eps, value=-lr)


def make_batch(R):
    g = torch.Generator().manual_seed(11)
    o, d, rgb = torch.randn(R, 3, generator=g), torch.randn(R, 3, generator=g), torch.rand(R, 3, generator=g)
    return o, d, rgb, torch.zeros(R, 1), torch.ones(R, 1)


def charb(pr, gt):
    return torch.sqrt(((gt - pr) ** 2).sum(-1) + 0.001)


def run_steps(world, rank, R, steps=3, nz=False):
    from nu_nerf_b200 import dist as nd
    net = TinyField(nz)
    tr = nd.DataParallelTrainer(net, net.render, charb, adam_fn=torch_adam, lr_fn=lambda s: 1e-2, normal_ori=nz)
    o, d, rgb, near, far = make_batch(R)
    idx = nd.shard_batch(torch.arange(R), rank, world)
    losses = []
    for s in range(steps):
        loss = tr.step(o[idx], d[idx], rgb[idx], near[idx], far[idx], (15000 if nz else 10000) + s)
        if world > 1:
            dist.all_reduce(loss)
        losses.append(loss.item())
    return tr.fp.flat.clone(), tr.fp.grad.clone(), losses


def _worker(rank, world, port, R, q, nz=False):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        flat, grad, losses = run_steps(world, rank, R, nz=nz)
        q.put((rank, flat.numpy(), grad.numpy(), losses))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(300)
@pytest.mark.parametrize("nz", [False, True])
def test_two_gloo_ranks_equal_the_single_process_large_batch_step(nz):
    """nz: with the outputs / loss list of the non-zero-thickness stage-1 configs at step >= 15000 -- the normal-orientation
    mean over the global ray count and the outer-regularisation mean over the candidate rays of ALL ranks (a data-dependent
    subset per rank: one scalar all-reduce of the counts)."""
    R, world = 64, 2
    ref_flat, ref_grad, ref_losses = run_steps(1, 0, R, nz=nz)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, R, q, nz)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, flat, grad, losses in res:
        assert torch.allclose(torch.from_numpy(grad), ref_grad, rtol=1e-5, atol=1e-7), f"rank {rank}: gradient differs"
        assert torch.allclose(torch.from_numpy(flat), ref_flat, rtol=1e-5, atol=1e-7), f"rank {rank}: weights differ"
        assert all(abs(a - b) < 1e-5 for a, b in zip(losses, ref_losses))


def test_shard_batch_is_a_partition():
    from nu_nerf_b200 import dist as nd
    idx = nd.batch_indices(1000, 3, 64)
    parts = [nd.shard_batch(idx, r, 4) for r in range(4)]
    assert sorted(torch.cat(parts).tolist()) == idx.tolist()
    assert all(p.numel() == 16 for p in parts)


def test_flat_parameters_alias_module_parameters():
    from nu_nerf_b200 import dist as nd
    net = TinyField()
    before = {k: v.clone() for k, v in net.state_dict().items()}
    fp = nd.FlatParameters(net)
    for k, v in net.state_dict().items():
        assert torch.equal(v, before[k])
    fp.flat.add_(1.0)
    for k, v in net.state_dict().items():
        assert torch.equal(v, before[k] + 1.0)
    assert all(p.grad.data_ptr() >= fp.grad.data_ptr() for p in net.parameters())

"""Device-side ray feeder (nu_nerf_b200/feeder.py) against the UNMODIFIED reference's ray-table code
(tests/golden/feeder.npz, made by tests/golden/make_golden_feeder.py): ray construction for both dataset conventions,
the shuffle (same CPU generator draw as ZT:195), training slices, reshuffle rule and the data-parallel sharding.
Runs on CPU tensors (the feeder is device-agnostic torch indexing); on the GPU box the same code keeps the table in HBM."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN


def _g():
    G = np.load(os.path.join(GOLDEN, "feeder.npz"))
    return {k: torch.from_numpy(G[k]) for k in G.files}


def test_nerf_ray_table_matches_reference():
    from nu_nerf_b200 import feeder
    G = _g()
    batch, rn, h, w = feeder.construct_nerf_ray_batch(G["imgs"], G["Ks"], G["poses"], G["masks"])
    assert (rn, h, w) == (60, 4, 5)
    for k in ("rgbs", "idxs", "masks"):
        assert torch.equal(batch[k], G["nerf_" + k]), k
    for k in ("rays_o", "rays_d"):
        assert (batch[k] - G["nerf_" + k]).abs().max().item() < 1e-6, k


def test_plain_ray_table_and_world_rays_match_reference():
    from nu_nerf_b200 import feeder
    G = _g()
    batch, rn, _, _ = feeder.construct_ray_batch(G["imgs"], G["Ks"])
    assert torch.equal(batch["idxs"], G["plain_idxs"]) and torch.equal(batch["rgbs"], G["plain_rgbs"])
    assert (batch["dirs"] - G["plain_dirs"]).abs().max().item() < 1e-6
    ro, rd = feeder.world_rays(batch["dirs"], batch["idxs"], G["poses"])
    assert (ro - G["plain_rays_o"]).abs().max().item() < 1e-6
    assert (rd - G["plain_rays_d"]).abs().max().item() < 1e-6


def test_shuffle_and_slices_match_reference():
    from nu_nerf_b200 import feeder
    G = _g()
    batch, rn, _, _ = feeder.construct_nerf_ray_batch(G["imgs"], G["Ks"], G["poses"], G["masks"])
    f = feeder.DeviceRayFeeder(batch, seed=7, perm_device="cpu")       # the reference's torch.manual_seed(7); randperm
    assert torch.equal(f.perm, G["perm"])
    b0 = f(0, 8)
    b1 = f(1, 8)
    assert torch.equal(b0["rays_o"], G["slice0_rays_o"]) and torch.equal(b1["rgbs"], G["slice1_rgbs"])
    assert set(b0) == {"rgbs", "idxs", "rays_o", "rays_d", "masks"}


def test_reshuffle_rule_and_rank_sharding():
    from nu_nerf_b200 import feeder
    G = _g()
    batch, rn, _, _ = feeder.construct_nerf_ray_batch(G["imgs"], G["Ks"], G["poses"])
    f = feeder.DeviceRayFeeder(batch, seed=3)
    first = f.perm.clone()
    for s in range(6):                      # 6 x 8 = 48 rays consumed; 48 + 8 < 60: no reshuffle yet
        f(s, 8)
    assert torch.equal(f.perm, first) and f.i == 48
    f(6, 8)                                 # 56 + 8 >= 60 -> reshuffled (ZT:452)
    assert f.i == 0 and not torch.equal(f.perm, first)
    # two ranks see disjoint, interleaved halves of the single-process batch
    one = feeder.DeviceRayFeeder(batch, seed=5)(0, 16)
    r0 = feeder.DeviceRayFeeder(batch, rank=0, world=2, seed=5)(0, 8)
    r1 = feeder.DeviceRayFeeder(batch, rank=1, world=2, seed=5)(0, 8)
    assert torch.equal(one["rays_d"][0::2], r0["rays_d"]) and torch.equal(one["rays_d"][1::2], r1["rays_d"])


def test_product_synthetic_inputs_equal_the_oracles():
    """bench.py's product arm draws its rays from nu_nerf_b200.synthetic; the oracle / goldens use their own copy."""
    from nu_nerf_b200 import synthetic as syn
    from oracle import nunerf_oracle as orc
    for a, b in zip(syn.synthetic_rays(33) + syn.synthetic_uniforms(33) + (syn.synthetic_targets(33),),
                    orc.synthetic_rays(33) + orc.synthetic_uniforms(33) + (orc.synthetic_targets(33),)):
        assert torch.equal(a, b)


def test_image_eval_source_views():
    """feeder.image_eval_source (the per-view source of test_step, ZT:397-411): the rays of view i are exactly the rows of
    the all-view ray table that belong to image i, for both dataset conventions."""
    from nu_nerf_b200 import feeder
    g = torch.Generator().manual_seed(4)
    imn, h, w = 3, 5, 7
    imgs = torch.rand(imn, 3, h, w, generator=g)
    K = torch.tensor([[20.0, 0, 3.5], [0, 20.0, 2.5], [0, 0, 1]])[None].repeat(imn, 1, 1)
    poses = torch.cat([torch.linalg.qr(torch.randn(imn, 3, 3, generator=g))[0], torch.randn(imn, 3, 1, generator=g)], -1)
    depth, mask = torch.rand(imn, h, w, generator=g), torch.rand(imn, h, w, generator=g) > 0.5
    full, rn, _, _ = feeder.construct_nerf_ray_batch(imgs, K, poses)
    src = feeder.image_eval_source(imgs, K, poses, is_nerf=True, depths=depth, masks=mask)
    for i in range(imn):
        v = src(i)
        rows = slice(i * h * w, (i + 1) * h * w)
        assert (v["h"], v["w"]) == (h, w) and v["gt_mask"].dtype == torch.int32
        assert torch.equal(v["rays_o"], full["rays_o"][rows]) and torch.equal(v["rays_d"], full["rays_d"][rows])
        assert torch.equal(v["rgbs"], full["rgbs"][rows]) and torch.equal(v["gt_depth"], depth[i])
    full2, _, _, _ = feeder.construct_ray_batch(imgs, K)
    src2 = feeder.image_eval_source(imgs, K, poses, is_nerf=False)
    for i in range(imn):
        v = src2(i)
        rows = slice(i * h * w, (i + 1) * h * w)
        ro, rd = feeder.world_rays(full2["dirs"][rows], full2["idxs"][rows], poses)
        assert torch.allclose(v["rays_o"], ro, atol=1e-6) and torch.allclose(v["rays_d"], rd, atol=1e-6)
        assert "gt_depth" not in v



class _FakeDatabase:
    """The subset of the reference's BaseDatabase interface the renderer reads (dataset/database.py): uint8 images,
    [3,4] poses, [3,3] intrinsics, (depth, mask)."""

    def __init__(self, n, h, w, seed=6):
        g = np.random.default_rng(seed)
        self.imgs = g.integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)
        self.K = np.array([[20.0, 0, w / 2], [0, 20.0, h / 2], [0, 0, 1]], np.float32)
        q = np.linalg.qr(g.standard_normal((n, 3, 3)))[0]
        self.poses = np.concatenate([q, g.standard_normal((n, 3, 1))], -1).astype(np.float32)
        self.depth = g.random((n, h, w)).astype(np.float32)
        self.mask = g.random((n, h, w)) > 0.5

    def get_image(self, i): return self.imgs[i]
    def get_pose(self, i): return self.poses[i]
    def get_K(self, i): return self.K
    def get_depth(self, i): return self.depth[i], self.mask[i]


@pytest.mark.parametrize("is_nerf", [True, False])
def test_attach_database_builds_the_ray_sources(is_nerf):
    """NeROShapeRenderer.attach_database (the dataset half of _init_dataset, ZT:167-197): a BaseDatabase-like object
    becomes the shuffled train ray source and the per-view eval source -- every served ray is a row of the all-view ray
    table with its own pixel colour, and batches do not repeat rays within an epoch."""
    from nu_nerf_b200 import feeder
    from nu_nerf_b200.renderer_zerothick import NeROShapeRenderer, load_default_cfg
    n, h, w = 4, 5, 6
    db = _FakeDatabase(n, h, w)
    cfg = load_default_cfg()
    cfg["database_name"] = None                      # no reference database package here: attach explicitly
    cfg["is_nerf"] = is_nerf
    torch.manual_seed(0)
    net = NeROShapeRenderer(cfg, training=True)
    assert net.ray_source is None
    net.attach_database(db, train_ids=[0, 1, 2], test_ids=[3])
    imgs = torch.from_numpy(db.imgs[:3].astype(np.float32) / 255.0).permute(0, 3, 1, 2)
    Ks, poses = torch.from_numpy(np.stack([db.K] * 3)), torch.from_numpy(db.poses[:3])
    if is_nerf:
        full, rn, _, _ = feeder.construct_nerf_ray_batch(imgs, Ks, poses)
        table_d = torch.nn.functional.normalize(full["rays_d"], dim=-1)
    else:
        full, rn, _, _ = feeder.construct_ray_batch(imgs, Ks)
        o_, d_ = feeder.world_rays(full["dirs"], full["idxs"], poses)
        full["rays_o"], table_d = o_, d_
    seen = []
    for step in range(3):
        b = net.ray_source(step, 20)
        assert b["rays_o"].shape == (20, 3) and b["rgbs"].shape == (20, 3)
        d = torch.nn.functional.normalize(b["rays_d"], dim=-1)
        for r in range(20):
            m = ((full["rays_o"] - b["rays_o"][r]).abs().sum(-1) < 1e-5) & ((table_d - d[r]).abs().sum(-1) < 1e-5)
            idx = m.nonzero().flatten()
            assert idx.numel() >= 1, (step, r)
            assert any(torch.allclose(full["rgbs"][i], b["rgbs"][r]) for i in idx.tolist())
            seen.append(int(idx[0]))
    assert len(set(seen)) == len(seen)               # one epoch: no ray twice
    v = net.eval_source(0)
    assert (v["h"], v["w"]) == (h, w) and v["rays_o"].shape == (h * w, 3)
    assert torch.allclose(v["rgbs"], torch.from_numpy(db.imgs[3].astype(np.float32) / 255.0).reshape(-1, 3))
    assert torch.equal(v["gt_depth"], torch.from_numpy(db.depth[3]))

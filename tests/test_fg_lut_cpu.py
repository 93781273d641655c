"""The split-sum table: the product's generator against the reference's asset (committed as tests/golden/
fg_lut_reference.npz by tests/golden/make_golden.py), and the product's search order (field.py:583 convention)."""
import os

import numpy as np


def test_generator_reproduces_the_reference_asset_within_its_sampling_noise(reference_fg_lut):
    from nu_nerf_b200.fg_lut import make_fg_lut
    ours, ref = make_fg_lut(), reference_fg_lut
    assert ours.shape == ref.shape == (1, 256, 256, 2) and ours.dtype == np.float32
    err = np.abs(ours - ref)
    # same integral (height-correlated Smith GGX split sum): the residual is Monte-Carlo noise of 1024 samples per texel,
    # largest at grazing NoV; it shrinks to 4.4e-3 / 2e-4 at 8192 samples (checked when the generator was written)
    assert err.max() < 1.5e-2, err.max()
    assert err.mean() < 1e-3, err.mean()
    assert np.abs(ours[0, :, -1] - ref[0, :, -1]).max() < 2e-3      # NoV -> 1 column


def test_search_order_follows_the_reference_convention(reference_fg_lut, tmp_path, monkeypatch):
    from nu_nerf_b200 import fg_lut
    # 1. $NUNERF_FG_LUT (set for the whole session by conftest.reference_fg_lut)
    assert np.array_equal(fg_lut.load_fg_lut(), reference_fg_lut)
    # 2. assets/bsdf_256_256.bin relative to the working directory, as the reference opens it
    monkeypatch.delenv("NUNERF_FG_LUT")
    (tmp_path / "assets").mkdir()
    marked = reference_fg_lut.copy()
    marked[0, 3, 5, 1] = 0.125
    marked.tofile(str(tmp_path / "assets" / "bsdf_256_256.bin"))
    monkeypatch.chdir(tmp_path)
    assert np.array_equal(fg_lut.load_fg_lut(), marked)
    # 3. nothing found: the generated table
    os.remove(str(tmp_path / "assets" / "bsdf_256_256.bin"))
    assert np.array_equal(fg_lut.load_fg_lut(), fg_lut.make_fg_lut())


def test_fresh_renderer_registers_the_loaded_table(reference_fg_lut):
    import torch
    from nu_nerf_b200.renderer_zerothick import NeROShapeRenderer, load_default_cfg
    net = NeROShapeRenderer(load_default_cfg(), training=False)
    assert torch.equal(net.color_network.FG_LUT, torch.from_numpy(reference_fg_lut))

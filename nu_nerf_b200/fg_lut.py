"""Split-sum environment-BRDF table (the "FG LUT" sampled by the shading network, field.py:583,719-723).

The reference ships the table as a binary asset, opens it as `assets/bsdf_256_256.bin` relative to the working
directory (field.py:583) and registers it as the buffer `color_network.FG_LUT`, so reference checkpoints carry their
own copy (which load_state_dict installs).  `load_fg_lut()` follows the same convention:

  1. $NUNERF_FG_LUT (a path to the 524 288-byte float32 file),
  2. `assets/bsdf_256_256.bin` relative to the working directory -- a run from the reference's tree gets its table,
  3. otherwise the table is integrated here (`make_fg_lut`): GGX importance sampling of the split-sum integral (Karis
     2013) with the HEIGHT-CORRELATED Smith visibility term, alpha = roughness^2, 1024 Hammersley points per texel,
     texel centres at (i + 1/2) / 256.  Against the reference asset this generator is within 1.3e-2 (max, at grazing
     NoV) / 8e-4 (mean) and converges to it with the sample count (4.4e-3 / 2e-4 at 8192 points): the asset is the same
     integral (tests/test_fg_lut_cpu.py holds the numbers against the committed fixture of the asset).

Layout [1, 256(roughness), 256(NoV), 2] so that texel (u=NoV, v=roughness) is [0, v, u].
"""
import functools
import os

import numpy as np

ASSET_RELPATH = os.path.join("assets", "bsdf_256_256.bin")


def _radical_inverse_vdc(n):
    bits = np.arange(n, dtype=np.uint64)
    bits = ((bits << np.uint64(16)) | (bits >> np.uint64(16))) & np.uint64(0xFFFFFFFF)
    bits = ((bits & np.uint64(0x55555555)) << np.uint64(1)) | ((bits & np.uint64(0xAAAAAAAA)) >> np.uint64(1))
    bits = ((bits & np.uint64(0x33333333)) << np.uint64(2)) | ((bits & np.uint64(0xCCCCCCCC)) >> np.uint64(2))
    bits = ((bits & np.uint64(0x0F0F0F0F)) << np.uint64(4)) | ((bits & np.uint64(0xF0F0F0F0)) >> np.uint64(4))
    bits = ((bits & np.uint64(0x00FF00FF)) << np.uint64(8)) | ((bits & np.uint64(0xFF00FF00)) >> np.uint64(8))
    return bits.astype(np.float64) * 2.3283064365386963e-10


@functools.lru_cache(maxsize=2)
def make_fg_lut(res=256, n_samples=1024):
    """Returns float32 [1, res, res, 2] (scale, bias) of the split-sum integral."""
    f32 = np.float32
    i = np.arange(n_samples)
    xi1 = ((i + 0.5) / n_samples).astype(f32)[None, None, :]
    xi2 = _radical_inverse_vdc(n_samples).astype(f32)[None, None, :]
    cphi = np.cos(2.0 * np.pi * xi1).astype(f32)
    nov = ((np.arange(res) + 0.5) / res).astype(f32)[None, :, None]
    vx, vz = np.sqrt(1.0 - nov * nov), nov
    out = np.empty((res, res, 2), f32)
    chunk = 16
    for r0 in range(0, res, chunk):                    # rows of constant roughness, a few at a time (memory)
        rough = ((np.arange(r0, min(r0 + chunk, res)) + 0.5) / res).astype(f32)[:, None, None]
        a2 = (rough * rough) ** 2
        cos_t = np.sqrt((1.0 - xi2) / (1.0 + (a2 - 1.0) * xi2))
        sin_t = np.sqrt(np.maximum(0.0, 1.0 - cos_t * cos_t))
        hx, hz = sin_t * cphi, cos_t
        voh = np.maximum(vx * hx + vz * hz, 0.0)
        nol = np.maximum(2.0 * voh * hz - vz, 0.0)
        lam = lambda c: (-1.0 + np.sqrt(1.0 + a2 * (1.0 - c * c) / np.maximum(c * c, 1e-12))) * 0.5
        g = 1.0 / (1.0 + lam(nol) + lam(nov))           # height-correlated Smith masking-shadowing
        gvis = np.where(nol > 0, g * voh / np.maximum(hz * nov, 1e-8), 0.0)
        fc = (1.0 - voh) ** 5
        out[r0:r0 + chunk, :, 0] = ((1.0 - fc) * gvis).mean(-1)
        out[r0:r0 + chunk, :, 1] = (fc * gvis).mean(-1)
    return out[None].astype(np.float32)


def load_fg_lut(res=256):
    """The table a freshly constructed shading network registers (see the module docstring for the search order)."""
    for path in (os.environ.get("NUNERF_FG_LUT"), ASSET_RELPATH):
        if path and os.path.isfile(path) and os.path.getsize(path) == res * res * 2 * 4:
            return np.fromfile(path, dtype=np.float32).reshape(1, res, res, 2).copy()
    return make_fg_lut(res).copy()

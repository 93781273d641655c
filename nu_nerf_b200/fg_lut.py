"""Split-sum environment-BRDF table (the "FG LUT" sampled by the shading network, field.py:583,719-723).

The reference ships the table as a binary asset and registers it as the buffer `color_network.FG_LUT`
(so reference checkpoints carry their own copy, which load_state_dict installs).  For a freshly
constructed renderer we integrate the table ourselves: GGX importance sampling of the split-sum
integral (Karis 2013), Smith-Schlick visibility with k = alpha/2, alpha = roughness^2, Hammersley
points.  Layout [1, 256(roughness), 256(NoV), 2] so that texel (u=NoV, v=roughness) is [0, v, u].
"""
import functools

import numpy as np


def _radical_inverse_vdc(n):
    bits = np.arange(n, dtype=np.uint64)
    bits = ((bits << np.uint64(16)) | (bits >> np.uint64(16))) & np.uint64(0xFFFFFFFF)
    bits = ((bits & np.uint64(0x55555555)) << np.uint64(1)) | ((bits & np.uint64(0xAAAAAAAA)) >> np.uint64(1))
    bits = ((bits & np.uint64(0x33333333)) << np.uint64(2)) | ((bits & np.uint64(0xCCCCCCCC)) >> np.uint64(2))
    bits = ((bits & np.uint64(0x0F0F0F0F)) << np.uint64(4)) | ((bits & np.uint64(0xF0F0F0F0)) >> np.uint64(4))
    bits = ((bits & np.uint64(0x00FF00FF)) << np.uint64(8)) | ((bits & np.uint64(0xFF00FF00)) >> np.uint64(8))
    return bits.astype(np.float64) * 2.3283064365386963e-10


@functools.lru_cache(maxsize=2)
def make_fg_lut(res=256, n_samples=512):
    """Returns float32 [1, res, res, 2] (scale, bias) of the split-sum integral."""
    i = np.arange(n_samples)
    xi1 = ((i + 0.5) / n_samples)[None, None, :]
    xi2 = _radical_inverse_vdc(n_samples)[None, None, :]
    nov = ((np.arange(res) + 0.5) / res)[None, :, None]
    rough = ((np.arange(res) + 0.5) / res)[:, None, None]
    a = rough * rough
    phi = 2.0 * np.pi * xi1
    cos_t = np.sqrt((1.0 - xi2) / (1.0 + (a * a - 1.0) * xi2))
    sin_t = np.sqrt(np.maximum(0.0, 1.0 - cos_t * cos_t))
    hx, hz = sin_t * np.cos(phi), cos_t
    vx, vz = np.sqrt(1.0 - nov * nov), nov
    voh = np.maximum(vx * hx + vz * hz, 0.0)
    lz = 2.0 * voh * hz - vz
    nol, noh = np.maximum(lz, 0.0), np.maximum(hz, 0.0)
    k = a / 2.0
    g = (nol / (nol * (1 - k) + k)) * (nov / (nov * (1 - k) + k))
    gvis = np.where(nol > 0, g * voh / np.maximum(noh * nov, 1e-8), 0.0)
    fc = (1.0 - voh) ** 5
    A = ((1.0 - fc) * gvis).mean(-1)
    B = (fc * gvis).mean(-1)
    return np.stack([A, B], -1)[None].astype(np.float32)

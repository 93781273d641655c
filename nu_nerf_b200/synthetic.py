"""Seeded synthetic inputs of SURVEY 8(d) for benchmarks and tools (no dataset, no network access): rays on a radius-3
sphere looking at the origin with 0.3 jitter, the two uniform draws of sample_ray, random target colours.  Pure torch CPU
generators, so every rank / run draws the same values (the test oracle keeps its own copy of these three functions)."""
import torch
import torch.nn.functional as F


def synthetic_rays(R, seed=1):
    g = torch.Generator().manual_seed(seed)
    o = 3.0 * F.normalize(torch.randn(R, 3, generator=g), dim=-1)
    d = F.normalize(-o + 0.3 * torch.randn(R, 3, generator=g), dim=-1)
    return o, d


def synthetic_uniforms(R, seed=2):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(R, 1, generator=g), torch.rand(R, 32, generator=g)


def synthetic_targets(R, seed=3):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(R, 3, generator=g)

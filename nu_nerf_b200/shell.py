"""Non-zero-thickness bounce geometry of network/renderer.py:1690-2009, per hit ray and in closed form.

The outer mesh is one face of a thin glass shell of thickness tau (ThicknessNetwork(x) * 0.01).  At a hit the shell is
modelled locally as the gap between two concentric spheres of radius r = 1 / sqrt(|K|) (K = the Gaussian curvature of the
mesh interpolated at the hit) and r -+ tau: the ray refracts into the glass (ratio eta), crosses it over the chord length
|r cos(theta_t) - sqrt((r cos(theta_t))^2 -+ 2 r tau + tau^2)|, and refracts again at the second sphere with the ratio
`eta_other` = eta_inner / eta.  Leaving the object the same construction runs backwards: the recorded hit lies on the
OUTER face, so the hit point is first pulled back along the ray onto the inner face (with a new normal from the sphere
centre) before the two refractions.

`shell_bounce` is a plain, differentiable torch function of per-ray tensors (any device): the renderer evaluates it on the
GPU on the [M, .] tensors of the rays that hit (M <= rays per bounce), autograd carries the loss to the IoR and thickness
networks through it.  Every expression keeps the reference's operation order (fp32 rounding included), the two curvature-
sign branches are folded into one expression with an exact sign factor (a - b == a + (-1) * b bit for bit).
"""
import torch
import torch.nn.functional as F

IOR_INNER = 1.0 / 1.0001          # NZ:1739 -- "we now assume inner is air": 1 / 1.0001 + 0 * f(IoRint_pred)


def _unit(v):
    return v / (torch.linalg.norm(v, dim=-1, keepdim=True) + 0.0001)


def signed_normal(n_raw, inside):
    """NZ:1648, :1659: unit mesh normal, pointing against the ray that travels inside the object."""
    n = F.normalize(n_raw, dim=-1)
    return -n if inside else n


def shell_bounce(x, normal, d, g_k, ior_sig, thick_sig, inside):
    """One bounce of NZ:1690-2009 on the M rays that hit the mesh.
      x [M,3] hit point, normal [M,3] signed unit mesh normal (signed_normal), d [M,3] incoming direction,
      g_k [M,1] interpolated Gaussian curvature, ior_sig / thick_sig [M,1] sigmoid outputs of IORs_pred / thickness_pred,
      inside: the ray travels inside the object (leaving it).
    Returns a dict:
      ok [M] bool          the first refraction is not a total internal reflection (`converged_out`, NZ:1768)
      ok_idx [K]           its row indices (ONE host sync); the tensors below marked [K,.] are over these rows
      tir [M] bool         ok and neither later refraction of the shell was clamped (`tir`, NZ:1773, :1860, :1939, :1993)
      x_mod [M,3]          the hit point (pulled back onto the inner face when leaving, NZ:1918, :1930)
      normal [K,3]         signed unit mesh normal (`gradient_mesh`)
      ratio [K,1]          the effective IoR ratio of the first refraction (`ior_ratios`)
      start, dir [K,3]     origin and direction of the next segment."""
    cos_i = torch.sum(normal * -d, dim=-1, keepdim=True)
    sin2_i = 1 - (cos_i * cos_i)
    ior = 1 / (ior_sig * 1.0 + 0.6)                                       # NZ:1733-1734
    ior_inner = torch.full_like(ior, IOR_INNER)
    ior_other = ior_inner / ior                                           # NZ:1744
    thick = thick_sig * 0.01                                              # NZ:1746-1747
    if inside:                                                            # NZ:1753-1756
        ior, ior_other = 1 / ior_other, 1 / ior
    ok = ~(ior * ior * sin2_i > 0.999).flatten()                          # NZ:1768
    ok_idx = ok.nonzero().squeeze(1)
    sel = lambda t: t.index_select(0, ok_idx)
    ior, ior_other, thick = sel(ior), sel(ior_other), sel(thick)
    sin2_t = sel(sin2_i) * ior * ior
    gk = sel(g_k)
    r = 1 / torch.sqrt(torch.clamp(torch.abs(gk), min=0.000001))          # NZ:1790-1791
    r = torch.nan_to_num(r, 0.1)
    x_k, n_k, d_k, cos_k = sel(x), sel(normal), sel(d), sel(cos_i)
    cos_t = torch.sqrt(torch.clamp(1 - sin2_t, min=0.0001))
    x_mod = x
    tir_k = torch.ones_like(ok_idx, dtype=torch.bool)
    if not inside:
        # s = -1 where the surface is convex towards the ray (K >= 0), +1 where concave
        s = torch.where(gk >= 0, -1.0, 1.0)
        d_in = _unit(ior * d_k + (ior * cos_k - torch.sqrt(torch.clamp(1 - sin2_t, min=0.0001))) * n_k)   # NZ:1812-1813
        n_in, x_in = n_k, x_k
    else:
        s = torch.where(gk <= 0, -1.0, 1.0)                                # NZ:1864 (the normal is flipped inside)
        # pull the hit back onto the inner face of the shell (NZ:1884-1930)
        c_r = r * cos_k
        delta = torch.sqrt(torch.clamp(c_r * c_r + s * (2 * r * thick) + thick * thick, min=0.0001))
        length = torch.abs(c_r - delta)
        center = x_k + s * (n_k * r)
        x_in = x_k - length * d_k
        n_mod = -s * (x_in - center)
        n_in = _unit(n_mod)                                                # NZ:1933
        x_mod = x.index_copy(0, ok_idx, x_in)
        cos_m = torch.sum(n_in * -d_k, dim=-1, keepdim=True)
        sin2_m = 1 - cos_m * cos_m
        e2 = sin2_m * ior * ior
        tir_k = tir_k & ~(e2 > 0.999).flatten()                            # NZ:1939
        sin2_tm = torch.clamp(e2, max=0.999)
        d_in = _unit(ior * d_k + (ior * cos_m - torch.sqrt(torch.clamp(1 - sin2_tm, min=0.0001))) * n_in)   # NZ:1941-1942
    # cross the shell: chord of the concentric sphere of radius r -+ tau (NZ:1816-1847, :1946-1987)
    ctr = r * cos_t
    delta = torch.sqrt(torch.clamp(ctr * ctr + s * (2 * r * thick) + thick * thick, min=0.0001))
    length = torch.abs(ctr - delta)
    center = x_in + s * (n_in * r)
    start = x_in + d_in.reshape(-1, 3) * (length.reshape(-1, 1) + 0.001)
    n_after = _unit(-s * (start - center))
    cos_2 = torch.sum(n_after * -d_in, dim=-1, keepdim=True)
    sin2_2 = 1 - (cos_2 * cos_2)
    e3 = sin2_2 * ior_other * ior_other
    tir_k = tir_k & ~(e3 > 0.999).flatten()                                # NZ:1860, :1993
    sin2_t2 = torch.clamp(e3, max=0.999)
    d_next = _unit(ior_other * d_in + (ior_other * cos_2 - torch.sqrt(torch.clamp(1 - sin2_t2, min=0.0001))) * n_after)
    tir = ok.clone()
    tir[ok_idx] = tir_k
    return {"ok": ok, "ok_idx": ok_idx, "tir": tir, "x_mod": x_mod, "normal": n_k, "ratio": ior, "start": start,
            "dir": d_next}


def outside_depths(device):
    """The 64 inverse-depth sample distances of a ray that leaves the scene (NZ:2140-2143)."""
    z = torch.linspace(1e-3, 1.0 - 1.0 / (64 + 1.0), 64, device=device)
    return 1.0 / torch.flip(z, dims=[-1]) + 1.0 / 64

"""Size-independent properties at BASELINE.json's FULL sizes (32 768 rays x 160 samples; 1 Mi rays against the 99 904-
triangle outer mesh; 512^3 grid), where the CPU oracle would take minutes: per-ray independence (a batch equals the
concatenation of its chunks, bit for bit), linearity of the compositing in the colours, the identity that ties the
forward weights to the backward colour gradient, conservation (weights + final transmittance), sortedness of the sampled
depths, consistency of the compaction map, analytic hits on the sphere mesh, Euler characteristic of the extracted
surface.  The small-size parity against the oracle / the reference goldens is in the other test files."""
import numpy as np
import pytest
import torch

from conftest import uv_sphere

pytestmark = pytest.mark.gpu
DEV = "cuda"
R_FULL, S = 32768, 160


def _renderer(precision="bf16"):
    from nu_nerf_b200.renderer_zerothick import NeROShapeRenderer, load_default_cfg
    torch.manual_seed(0)
    cfg = load_default_cfg()
    cfg["precision"] = precision
    return NeROShapeRenderer(cfg, training=False).cuda()


@pytest.fixture(scope="module")
def sampled():
    """sample_ray on the full config-3 batch (SURVEY 8d rays / uniforms)."""
    from nu_nerf_b200 import synthetic as syn
    net = _renderer()
    o, d = (t.cuda() for t in syn.synthetic_rays(R_FULL))
    U0, U1 = (t.cuda() for t in syn.synthetic_uniforms(R_FULL))
    near, far = torch.full((R_FULL, 1), 0.8, device=DEV), torch.full((R_FULL, 1), 4.5, device=DEV)
    with torch.no_grad():
        z = net.sample_ray(o, d, near, far, 1.0, uniforms=(U0, U1))
    return net, o, d, near, far, (U0, U1), z


def test_sample_ray_sorted_bounded_and_chunk_invariant(sampled):
    net, o, d, near, far, (U0, U1), z = sampled
    assert z.shape == (R_FULL, S) and torch.isfinite(z).all()
    zi, zo = z[:, :128], z[:, 128:]
    assert (zi[:, 1:] >= zi[:, :-1]).all()                     # every merge keeps the row sorted (ZT:556-570)
    assert (zo[:, 1:] > zo[:, :-1]).all() and (zo[:, 0] > far[:, 0]).all()      # inverse-depth strata beyond far
    lo = near[:, 0] - 1.0 / 64 - 1e-6                           # coarse jitter (U0 - 0.5) * 2 / 64
    assert (zi[:, 0] >= lo).all() and (zi[:, -1] <= far[:, 0] + 1.0 / 64 + 1e-6).all()
    # rays are independent: any chunk of the batch reproduces its rows (bf16 MLP tiles are row-independent too)
    sl = slice(12288, 12288 + 4096)
    with torch.no_grad():
        zc = net.sample_ray(o[sl], d[sl], near[sl], far[sl], 1.0, uniforms=(U0[sl], U1[sl]))
    assert torch.equal(zc, z[sl])


def _geometry_lean(o, d, z):
    from nu_nerf_b200 import _lib
    R = z.shape[0]
    f = lambda *s: torch.empty(*s, device=DEV)
    i32 = lambda *s: torch.zeros(*s, dtype=torch.int32, device=DEV)
    g = dict(counts=i32(2), scratch=i32(2 * R), pts_in=f(R * S, 3), dists_in=f(R * S), dirs_in=f(R * S, 3),
             pts_out=f(R * S, 3), dists_out=f(R * S), dirs_out=f(R * S, 3), ray_map=i32(R, 10))
    _lib.call("nunerf_render_geometry", o.data_ptr(), d.data_ptr(), z.data_ptr(), R, S, None, None, None,
              g["counts"].data_ptr(), g["scratch"].data_ptr(), g["pts_in"].data_ptr(), g["dists_in"].data_ptr(),
              g["dirs_in"].data_ptr(), None, g["pts_out"].data_ptr(), g["dists_out"].data_ptr(), g["dirs_out"].data_ptr(), None,
              g["ray_map"].data_ptr())
    return g


def test_geometry_compaction_map_is_consistent(sampled):
    _, o, d, _, _, _, z = sampled
    g = _geometry_lean(o, d, z.contiguous())
    n_in, n_out = (int(v) for v in g["counts"].tolist())
    assert n_in + n_out == R_FULL * S and 0 < n_in < R_FULL * S
    rm = g["ray_map"].long()
    pop = sum(((rm[:, 2 + k] & 0xFFFFFFFF).unsqueeze(1) >> torch.arange(32, device=DEV) & 1).sum(1) for k in range(5))
    in_off = torch.cumsum(pop, 0) - pop
    assert torch.equal(rm[:, 0], in_off) and int(pop.sum()) == n_in             # exclusive scan of the per-ray counts
    assert torch.equal(rm[:, 1], torch.arange(R_FULL, device=DEV) * S - in_off)
    # the lists hold exactly the inside / outside points, in ray order
    assert (g["pts_in"][:n_in].norm(dim=-1) <= 1.0 + 1e-6).all() and (g["pts_out"][:n_out].norm(dim=-1) >= 1.0 - 1e-6).all()
    ray_of_in = torch.repeat_interleave(torch.arange(R_FULL, device=DEV), pop)
    dn = torch.nn.functional.normalize(d, dim=-1)
    assert (g["dirs_in"][:n_in] - dn[ray_of_in]).abs().max().item() <= 2e-7
    # a point of the inner list lies on its ray: (p - o) x d = 0
    resid = torch.linalg.cross(g["pts_in"][:n_in] - o[ray_of_in], dn[ray_of_in]).norm(dim=-1)
    assert resid.max().item() < 1e-5


def _composite(g, a_in, c_in, a_out, c_out, R, is_nerf, want_w=False, grads=None):
    from nu_nerf_b200 import _lib
    f = lambda *s: torch.empty(*s, device=DEV)
    rgb, raw, acc, bk = f(R, 3), f(R, 3), f(R), f(R, 3)
    w = f(R, S) if want_w else None
    _lib.call("nunerf_composite_fwd", a_in.data_ptr(), c_in.data_ptr(), a_out.data_ptr(), c_out.data_ptr(), None, R, S,
              is_nerf, rgb.data_ptr(), raw.data_ptr(), acc.data_ptr(), bk.data_ptr(), w.data_ptr() if want_w else None,
              g["ray_map"].data_ptr())
    out = dict(rgb=rgb, raw=raw, acc=acc, bk=bk, w=w)
    if grads is not None:
        g_rgb, g_acc, g_bk = grads
        da_in, dc_in, da_out, dc_out = f(a_in.shape[0]), f(a_in.shape[0], 3), f(a_out.shape[0]), f(a_out.shape[0], 3)
        _lib.call("nunerf_composite_bwd", a_in.data_ptr(), c_in.data_ptr(), a_out.data_ptr(), c_out.data_ptr(), None, R, S,
                  is_nerf, raw.data_ptr(), g_rgb.data_ptr(), g_acc.data_ptr() if g_acc is not None else None,
                  g_bk.data_ptr() if g_bk is not None else None, da_in.data_ptr(), dc_in.data_ptr(), da_out.data_ptr(),
                  dc_out.data_ptr(), g["ray_map"].data_ptr())
        out.update(da_in=da_in, dc_in=dc_in, da_out=da_out, dc_out=dc_out)
    return out


def test_compositing_properties_at_full_size(sampled):
    _, o, d, _, _, _, z = sampled
    g = _geometry_lean(o, d, z.contiguous())
    n_in, n_out = (int(v) for v in g["counts"].tolist())
    gen = torch.Generator(device=DEV).manual_seed(11)
    a_in = torch.rand(n_in, device=DEV, generator=gen) ** 4
    a_out = torch.rand(n_out, device=DEV, generator=gen) ** 4 * 0.5
    c_in, c_out = torch.rand(n_in, 3, device=DEV, generator=gen) * 0.5, torch.rand(n_out, 3, device=DEV, generator=gen) * 0.5
    base = _composite(g, a_in, c_in, a_out, c_out, R_FULL, 0, want_w=True)
    w = base["w"]
    # conservation: acc = sum of the weights = 1 - final transmittance (up to the 1e-7 floors), within [0, 1]
    assert (w >= 0).all() and (base["acc"] - w.sum(1)).abs().max().item() < 2e-6
    assert base["acc"].min().item() >= 0.0 and base["acc"].max().item() <= 1.0 + 1e-5
    # constant colour: rgb_raw = acc * colour
    ones_in, ones_out = torch.ones_like(c_in), torch.ones_like(c_out)
    const = _composite(g, a_in, ones_in, a_out, ones_out, R_FULL, 0)
    assert (const["raw"] - base["acc"][:, None]).abs().max().item() < 2e-6
    # is_nerf adds the white background (1 - acc): constant colour composites to exactly 1 up to rounding
    white = _composite(g, a_in, ones_in, a_out, ones_out, R_FULL, 1)
    assert (white["raw"] - 1.0).abs().max().item() < 2e-6
    # linearity in the colours
    c2_in, c2_out = torch.rand(n_in, 3, device=DEV, generator=gen), torch.rand(n_out, 3, device=DEV, generator=gen)
    other = _composite(g, a_in, c2_in, a_out, c2_out, R_FULL, 0)
    both = _composite(g, a_in, c_in + c2_in, a_out, c_out + c2_out, R_FULL, 0)
    assert (both["raw"] - base["raw"] - other["raw"]).abs().max().item() < 3e-6
    assert (both["bk"] - base["bk"] - other["bk"]).abs().max().item() < 3e-6
    # the background-only composite never exceeds what the outer samples alone can give: it ignores the inner alphas
    zero_in = _composite(g, torch.zeros_like(a_in), c_in, a_out, c_out, R_FULL, 0)
    assert (zero_in["bk"] - base["bk"]).abs().max().item() < 2e-6
    # backward: d(sum g . rgb) / d colour_k = w_k g_ray  -- ties the recomputed transmittance to the forward weights
    g_rgb = torch.randn(R_FULL, 3, device=DEV, generator=gen)
    bw = _composite(g, a_in, c_in, a_out, c_out, R_FULL, 0, grads=(g_rgb, None, None))
    rm = g["ray_map"].long()
    bits = torch.cat([((rm[:, 2 + k] & 0xFFFFFFFF).unsqueeze(1) >> torch.arange(32, device=DEV)) & 1 for k in range(5)], 1).bool()
    expect = w[..., None] * g_rgb[:, None, :]                      # [R, S, 3]
    inside01 = (base["raw"] >= 0) & (base["raw"] <= 1)             # clamp passes the gradient on [0, 1]
    expect = expect * inside01[:, None, :]
    assert (bw["dc_in"] - expect[bits]).abs().max().item() < 2e-6
    assert (bw["dc_out"] - expect[~bits]).abs().max().item() < 2e-6
    # d(sum acc) / d alpha_k = T_k - (sum_{j>k} w_j) / (1 - a_k + 1e-7) >= ... : check through a directional derivative
    g_acc = torch.ones(R_FULL, device=DEV)
    ba = _composite(g, a_in, c_in, a_out, c_out, R_FULL, 0, grads=(torch.zeros(R_FULL, 3, device=DEV), g_acc, None))
    eps = 1e-4
    dir_in, dir_out = torch.rand(n_in, device=DEV, generator=gen), torch.rand(n_out, device=DEV, generator=gen)
    a_in64, a_out64 = a_in.double(), a_out.double()

    def acc64(ai, ao):          # fp64 torch restatement of acc on the dense layout (ZT:773-776)
        alpha = torch.zeros(R_FULL, S, dtype=torch.float64, device=DEV)
        alpha[bits] = ai
        alpha[~bits] = ao
        T = torch.cumprod(torch.cat([torch.ones(R_FULL, 1, dtype=torch.float64, device=DEV), 1 - alpha + 1e-7], 1), 1)[:, :-1]
        return (alpha * T).sum(1)
    fd = (acc64(a_in64 + eps * dir_in, a_out64 + eps * dir_out) - acc64(a_in64 - eps * dir_in, a_out64 - eps * dir_out)) / (2 * eps)
    da = torch.zeros(R_FULL, S, dtype=torch.float64, device=DEV)
    dd = torch.zeros(R_FULL, S, dtype=torch.float64, device=DEV)
    da[bits], da[~bits] = ba["da_in"].double(), ba["da_out"].double()
    dd[bits], dd[~bits] = dir_in.double(), dir_out.double()
    an = (da * dd).sum(1)
    assert ((an - fd).abs() / (1 + fd.abs())).max().item() < 2e-3
    # chunk invariance, bit for bit (a ray's result depends on its own runs only)
    sl = slice(8192, 8192 + 4096)
    gs = _geometry_lean(o[sl].contiguous(), d[sl].contiguous(), z[sl].contiguous())
    i0, i1 = int(rm[sl.start, 0]), int(rm[sl.stop, 0])
    o0, o1 = int(rm[sl.start, 1]), int(rm[sl.stop, 1])
    part = _composite(gs, a_in[i0:i1].clone(), c_in[i0:i1].clone(), a_out[o0:o1].clone(), c_out[o0:o1].clone(), 4096, 0)
    for k in ("raw", "acc", "bk"):
        assert torch.equal(part[k], base[k][sl]), k


def test_bvh_trace_analytic_sphere_1mi_rays():
    """1 Mi rays against the config-4 outer mesh (UV sphere, radius 0.6, 99 904 triangles): hits agree with the analytic
    sphere outside a thin band around the silhouette, hit distances within the facet sagitta, and the reported triangle
    contains the hit point."""
    from nu_nerf_b200.tracer import TriangleBVH
    V, Fc = uv_sphere(0.6, 224, 224)
    Vt, Ft = torch.from_numpy(V).float().cuda(), torch.from_numpy(Fc).int().cuda()
    assert Ft.shape[0] == 99904
    bvh = TriangleBVH(Vt, Ft)
    N = 1 << 20
    gen = torch.Generator(device=DEV).manual_seed(3)
    o = 3.0 * torch.nn.functional.normalize(torch.randn(N, 3, device=DEV, generator=gen), dim=-1)
    d = torch.nn.functional.normalize(-o + 0.3 * torch.randn(N, 3, device=DEV, generator=gen), dim=-1)
    hit, tri, t = bvh.trace(o, d, return_t=True)
    b = (o * d).sum(-1)
    disc = b * b - ((o * o).sum(-1) - 0.36)
    closest = torch.sqrt(torch.clamp((o * o).sum(-1) - b * b, min=0))        # distance of the ray to the centre
    sag = 0.6 * (1 - np.cos(np.pi / 224))                                      # the mesh is inscribed: facets sit inside by <= sag
    sure_hit, sure_miss = closest < 0.6 - 4 * sag, closest > 0.6 + 1e-6
    assert (hit[sure_hit] == 1).all() and (hit[sure_miss] == 0).all()
    assert (tri[hit == 0] == 10000000).all() and ((tri[hit == 1] >= 0) & (tri[hit == 1] < 99904)).all()
    h = sure_hit
    t_an = -b[h] - torch.sqrt(disc[h])
    assert (t[h] >= t_an - 1e-5).all()                                         # the inscribed mesh is hit after the sphere
    # the hit point lies in the shell between the facets' deepest points and the sphere, and the triangle contains it
    p = o[h] + d[h] * t[h, None]
    rad = p.norm(dim=-1)
    assert (rad <= 0.6 + 1e-5).all() and (rad >= 0.6 - 2 * sag - 1e-5).all()
    tv = bvh.tri_verts[tri[h].long()].view(-1, 3, 3).double()
    p = p.double()
    n = torch.linalg.cross(tv[:, 1] - tv[:, 0], tv[:, 2] - tv[:, 0])
    nh = n / n.norm(dim=-1, keepdim=True)
    # in-plane distance of the hit point to its triangle from the three edge functions (the pole triangles are needles
    # 2e-4 wide, so a barycentric tolerance would be meaningless), and distance to its plane
    sd = []
    for i in range(3):
        a, b_ = tv[:, i], tv[:, (i + 1) % 3]
        e = b_ - a
        sd.append((torch.linalg.cross(e, p - a) * nh).sum(-1) / e.norm(dim=-1))
    outside = torch.clamp(-torch.minimum(torch.minimum(sd[0], sd[1]), sd[2]), min=0)
    assert outside.max().item() < 3e-6, outside.max().item()
    off_plane = ((p - tv[:, 0]) * nh).sum(-1).abs()
    assert off_plane.max().item() < 2e-5, off_plane.max().item()     # fp32 Moeller-Trumbore t from 3 units away
    # determinism / chunk invariance of the ids
    hit2, tri2 = bvh.trace(o[:4096], d[:4096])
    assert torch.equal(tri2, tri[:4096]) and torch.equal(hit2, hit[:4096])


def test_grid_sweep_and_marching_cubes_512():
    """Config 5 at full size: the 512^3 sweep of the (bumpy-sphere) initial field and the surface extracted from it form
    ONE closed oriented 2-manifold: F = 2 V - 4, every edge shared by exactly two triangles."""
    from nu_nerf_b200.sweep import extract_fields, marching_cubes
    from oracle import mc_oracle as mco
    net = _renderer()
    u = extract_fields(-torch.ones(3), torch.ones(3), 512, net.sdf_network.sdf, return_device=True)
    assert u.shape == (512, 512, 512) and torch.isfinite(u).all()
    assert (u[0] == 1.0).all() and (u[:, :, -1] == 1.0).all()                  # outside the unit sphere: outside_val
    inside = float((u < 0).float().mean())
    assert 0.05 < inside < 0.09                                                # a sphere of radius ~0.5 in [-1, 1]^3: 6.5 %
    v, t = marching_cubes(u, 0.0)
    V, E, F, boundary, nonmanifold, inconsistent = mco.mesh_stats(t)
    assert (boundary, nonmanifold, inconsistent) == (0, 0, 0)
    assert V - E + F == 2 and F == 2 * V - 4 and V == len(v)

"""TEST INFRASTRUCTURE ONLY -- CPU oracle (torch fp32) for the NU-NeRF stage-1 hot path.

An independent restatement of the reference algorithm (file:line citations are into
/root/reference/, "ZT" = network/renderer_zerothick.py).  Everything is *functional* over a
state_dict with the reference's key names, so the same tensors can be fed to the reference
(in the build container), to this oracle and to the CUDA product.

Pinning: tests/test_oracle_golden.py checks this file against fixtures under tests/golden/ that
were produced by the UNMODIFIED reference (tests/golden/make_golden.py, run in the build container
through oracle/ref_harness.py).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import it; the product package never does.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

SQRT2 = math.sqrt(2.0)


# ----------------------------------------------------------------------------- encodings
def pos_enc(x, n_freq):
    """field.py:14-61 -- [x, sin(2^k x), cos(2^k x)]_{k<n_freq}, blocks of width d in x,y,z order."""
    out = [x]
    for k in range(n_freq):
        f = float(2.0 ** k)
        out.append(torch.sin(x * f))
        out.append(torch.cos(x * f))
    return torch.cat(out, -1)


def _gen_binom(a, k):
    return np.prod(a - np.arange(k)) / math.factorial(k)


def _assoc_legendre_coeff(l, m, k):
    return ((-1) ** m * 2 ** l * math.factorial(l) / math.factorial(k) / math.factorial(l - k - m)
            * _gen_binom(0.5 * (l + k + m - 1.0), l))


def _sph_harm_coeff(l, m, k):
    return math.sqrt((2.0 * l + 1.0) * math.factorial(l - m) / (4.0 * math.pi * math.factorial(l + m))) \
        * _assoc_legendre_coeff(l, m, k)


def ide_tables(deg_view=5):
    """utils/ref_utils.py:39-83 -- (m,l) list with l=2^i, m=0..l and the z-polynomial matrix [l_max+1, 36]."""
    ml = [(m, 2 ** i) for i in range(deg_view) for m in range(2 ** i + 1)]
    l_max = 2 ** (deg_view - 1)
    mat = np.zeros((l_max + 1, len(ml)))
    for i, (m, l) in enumerate(ml):
        for k in range(l - m + 1):
            mat[k, i] = _sph_harm_coeff(l, m, k)
    return ml, mat.astype(np.float32)


_ML, _MAT = ide_tables(5)
_MAT_T = torch.from_numpy(_MAT)
_SIGMA = torch.tensor([0.5 * l * (l + 1) for (_, l) in _ML], dtype=torch.float32)
_MS = [m for (m, _) in _ML]


def ide(xyz, kappa_inv):
    """utils/ref_utils.py:85-114 -- integrated directional encoding, 72 = cat(Re[36], Im[36])."""
    x, y, z = xyz[..., 0:1], xyz[..., 1:2], xyz[..., 2:3]
    vmz = torch.cat([z ** i for i in range(_MAT_T.shape[0])], -1)
    # (x+iy)^m by repeated complex multiplication
    re = [torch.ones_like(x)]
    im = [torch.zeros_like(x)]
    for _ in range(16):
        r, i = re[-1], im[-1]
        re.append(r * x - i * y)
        im.append(r * y + i * x)
    vre = torch.cat([re[m] for m in _MS], -1)
    vim = torch.cat([im[m] for m in _MS], -1)
    pz = vmz @ _MAT_T.to(xyz.dtype)
    att = torch.exp(-_SIGMA.to(xyz.dtype) * kappa_inv)
    return torch.cat([vre * pz * att, vim * pz * att], -1)


def linear_to_srgb(x):
    """utils/raw_utils.py:5-12."""
    eps = torch.finfo(torch.float32).eps
    return torch.where(x <= 0.0031308, 323.0 / 25.0 * x,
                       (211.0 * torch.clamp(x, min=eps) ** (5.0 / 12.0) - 11.0) / 200.0)


# ----------------------------------------------------------------------------- networks
# Test aid: emulate the ENGINE's tensor-core operand precision inside this fp32 oracle.  None = plain fp32 (the reference's
# arithmetic); "split" = operands carried as bf16 hi + lo planes (16 mantissa bits), "bf16" = one bf16 plane.  Both matmul
# operands and the gradient entering the matmul (the engine's dZ planes) are rounded, with straight-through derivatives.
# tests/test_engine_gpu.py uses it to show that the parameter gradients which differ from the fp32 reference by more than
# 1e-3 do so because of this operand rounding (ReLU / clamp decisions of a few samples flip), not because of the kernels.
OPERAND_PRECISION = None


def _rnd(x):
    if OPERAND_PRECISION is None or not x.dtype.is_floating_point:
        return x
    hi = x.to(torch.bfloat16).to(x.dtype)
    if OPERAND_PRECISION == "bf16":
        return hi
    return hi + (x - hi).to(torch.bfloat16).to(x.dtype)


class _Operand(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return _rnd(x)

    @staticmethod
    def backward(ctx, g):
        return g


def _mm(h, Wt, b=None):
    """h @ Wt (+ b), `Wt` already transposed to [in, out]."""
    if OPERAND_PRECISION is None:
        z = h @ Wt
    else:
        z = _Operand.apply(h) @ _Operand.apply(Wt)
    if b is not None:
        z = z + b
    if OPERAND_PRECISION is not None and z.requires_grad:
        z.register_hook(_rnd)
    return z


def wn_weight(sd, prefix):
    """nn.utils.weight_norm(dim=0): W = g * v / ||v||_row."""
    v, g = sd[prefix + ".weight_v"], sd[prefix + ".weight_g"]
    return g * v / v.norm(dim=1, keepdim=True)


def softplus100(x):
    return F.softplus(x, beta=100)


def sdf_forward(sd, x, prefix="sdf_network", with_grad=False):
    """SDFNetwork.forward (field.py:133-150) and, if with_grad, the explicit adjoint pass that equals
    autograd.grad(sdf, x) (field.py:158-170).  Returns out[N,257] (, grad[N,3])."""
    pe = pos_enc(x, 6)
    h = pe
    s_list, W_list = [], []
    for l in range(9):
        W = wn_weight(sd, f"{prefix}.lin{l}")
        b = sd[f"{prefix}.lin{l}.bias"]
        if l == 4:
            h = torch.cat([h, pe], -1) / SQRT2
        z = _mm(h, W.t(), b)
        W_list.append(W)
        if l < 8:
            h = softplus100(z)
            s_list.append(torch.sigmoid(100.0 * z))
        else:
            h = z
    if not with_grad:
        return h
    g = W_list[8][0:1, :].expand(x.shape[0], -1)              # d sdf / d a_7
    g_pe_skip = None
    for l in range(7, -1, -1):
        g = g * s_list[l]                                      # through softplus of layer l
        g = _mm(g, W_list[l])                                  # to the input of layer l
        if l == 4:
            g = g / SQRT2
            g_pe_skip = g[:, 217:]
            g = g[:, :217]
    g_pe = g + g_pe_skip
    # Jacobian of the positional encoding
    grad = g_pe[:, 0:3].clone()
    for k in range(6):
        f = float(2.0 ** k)
        grad = grad + g_pe[:, 3 + 6 * k: 6 + 6 * k] * (f * torch.cos(x * f))
        grad = grad - g_pe[:, 6 + 6 * k: 9 + 6 * k] * (f * torch.sin(x * f))
    return h, grad


def predictor(sd, prefix, x, act, exp_max=3.0):
    """make_predictor (field.py:371-408): 4 weight-normed layers, ReLU, output activation."""
    h = x
    for i, l in enumerate((0, 2, 4, 6)):
        h = _mm(h, wn_weight(sd, f"{prefix}.{l}").t(), sd[f"{prefix}.{l}.bias"])
        if i < 3:
            h = F.relu(h)
    if act == "sigmoid":
        return torch.sigmoid(h)
    if act == "exp":
        return torch.exp(torch.clamp(h, max=exp_max))
    return h


def nerfpp_forward(sd, pts4, views, prefix="outer_nerf"):
    """NeRFNetwork.forward (field.py:265-289), D=8, W=256, skip after layer 4, PE-10 / PE-4."""
    xpe = pos_enc(pts4, 10)
    vpe = pos_enc(views, 4)
    h = xpe
    for i in range(8):
        h = F.relu(_mm(h, sd[f"{prefix}.pts_linears.{i}.weight"].t(), sd[f"{prefix}.pts_linears.{i}.bias"]))
        if i == 4:
            h = torch.cat([xpe, h], -1)
    alpha = _mm(h, sd[f"{prefix}.alpha_linear.weight"].t(), sd[f"{prefix}.alpha_linear.bias"])
    feat = _mm(h, sd[f"{prefix}.feature_linear.weight"].t(), sd[f"{prefix}.feature_linear.bias"])
    h = torch.cat([feat, vpe], -1)
    h = F.relu(_mm(h, sd[f"{prefix}.views_linears.0.weight"].t(), sd[f"{prefix}.views_linears.0.bias"]))
    rgb = _mm(h, sd[f"{prefix}.rgb_linear.weight"].t(), sd[f"{prefix}.rgb_linear.bias"])
    return alpha, rgb


def fg_lookup(lut, u, v):
    """dr.texture(FG_LUT[1,256,256,2], uv, 'linear', 'clamp') (field.py:719-722): texel centres at
    (i+0.5)/256, u -> width axis, v -> height axis, clamp to edge."""
    H, W = lut.shape[1], lut.shape[2]
    fx = torch.clamp(u * W - 0.5, 0.0, W - 1.0)
    fy = torch.clamp(v * H - 0.5, 0.0, H - 1.0)
    x0 = torch.floor(fx)
    y0 = torch.floor(fy)
    tx, ty = fx - x0, fy - y0
    x0i, y0i = x0.long(), y0.long()
    x1i = torch.clamp(x0i + 1, max=W - 1)
    y1i = torch.clamp(y0i + 1, max=H - 1)
    t = lut[0]
    c00, c01 = t[y0i, x0i], t[y0i, x1i]
    c10, c11 = t[y1i, x0i], t[y1i, x1i]
    tx, ty = tx[..., None], ty[..., None]
    return (c00 * (1 - tx) + c01 * tx) * (1 - ty) + (c10 * (1 - tx) + c11 * tx) * ty


def shading_forward(sd, points, grads, view_dirs, feats, prefix="color_network", exp_max=3.0, extras=False):
    """AppShadingNetwork.forward (field.py:684-777) with human_light / sphere_direction off."""
    n = F.normalize(grads, dim=-1)
    v = F.normalize(view_dirs, dim=-1)
    nov = (n * v).sum(-1, keepdim=True)
    refl = nov * n * 2 - v
    x = torch.cat([feats, points], -1)
    metallic = predictor(sd, f"{prefix}.metallic_predictor", x, "sigmoid")
    rough = predictor(sd, f"{prefix}.roughness_predictor", x, "sigmoid")
    albedo = predictor(sd, f"{prefix}.albedo_predictor", x, "sigmoid")
    trans = predictor(sd, f"{prefix}.transmisstion_weight", x, "sigmoid")

    diffuse_albedo = (1 - metallic) * albedo
    diffuse_light = predictor(sd, f"{prefix}.outer_light", ide(n, torch.ones_like(rough)), "exp", exp_max)
    diffuse_color = diffuse_albedo * diffuse_light
    spec_albedo = 0.04 * (1 - metallic) + metallic * albedo

    ide_r = ide(refl, rough)
    ide_0 = ide(refl, torch.zeros_like(rough))
    ppe = pos_enc(points, 6)
    direct = predictor(sd, f"{prefix}.outer_light", ide_r, "exp", exp_max)
    direct0 = predictor(sd, f"{prefix}.outer_light", ide_0, "exp", exp_max)
    indirect = predictor(sd, f"{prefix}.inner_light", torch.cat([ppe, ide_r], -1), "exp", exp_max)
    indirect0 = predictor(sd, f"{prefix}.inner_light", torch.cat([ppe, ide_0], -1), "exp", exp_max)
    occ = predictor(sd, f"{prefix}.inner_weight", torch.cat([ppe.detach(), pos_enc(refl, 6).detach()], -1), "none")
    occ = occ * 0.5 + 0.5
    occ_c = torch.clamp(occ, 0.0, 1.0)
    light = indirect * occ_c + direct * (1 - occ_c)
    light0 = indirect0 * occ_c + direct0 * (1 - occ_c)

    t = torch.clamp(1 - nov, 0.0, 1.0)
    schlick = 0.04 + 0.96 * t * t * t * t * t
    refl_w = torch.clamp(schlick, 0.0, 1.0)
    refr = predictor(sd, f"{prefix}.refrac_light", torch.cat([ppe, pos_enc(v, 6)], -1), "exp", exp_max)
    fg = fg_lookup(sd[f"{prefix}.FG_LUT"], torch.clamp(nov[:, 0], 0.0, 1.0), torch.clamp(rough[:, 0], 0.0, 1.0))
    spec_ref = spec_albedo * fg[:, 0:1] + fg[:, 1:2]
    spec_color = spec_ref * light
    color = (diffuse_color + spec_color) * (1 - trans) + (refl_w * light0 + (1 - refl_w) * refr) * trans
    color = linear_to_srgb(color)
    info = {"reflective": refl, "occ_prob": occ, "transmission_weight": trans, "metallic": metallic}
    if extras:
        info.update({"roughness": rough, "albedo": albedo, "diffuse_light": diffuse_light, "light": light,
                     "light0": light0, "refraction_light": refr, "fg": fg, "nov": nov})
    return color, info


# ----------------------------------------------------------------------------- sampling
def near_far_from_sphere(o, d):
    """ZT:320-327."""
    a = (d * d).sum(-1, keepdim=True)
    b = 2.0 * (o * d).sum(-1, keepdim=True)
    mid = 0.5 * (-b) / a
    return torch.clamp(mid - 1.0, min=1e-3), mid + 1.0


def coarse_samples(near, far, U0, U1, n_samples=64, n_bg=32, perturb=True):
    """ZT:580-594 -- 64 stratified-shifted samples and 32 inverse-depth background samples."""
    t = torch.linspace(0.0, 1.0, n_samples)
    z = near + (far - near) * t[None, :]
    b = torch.linspace(1e-3, 1.0 - 1.0 / (n_bg + 1.0), n_bg)
    if perturb:
        z = z + (U0 - 0.5) * 2.0 / n_samples
        mids = 0.5 * (b[1:] + b[:-1])
        upper = torch.cat([mids, b[-1:]])
        lower = torch.cat([b[:1], mids])
        b = lower[None, :] + (upper - lower)[None, :] * U1
    else:
        b = b[None, :].expand(near.shape[0], -1)
    z_out = far / torch.flip(b, dims=[-1]) + 1.0 / n_bg
    return z, z_out


def sample_pdf_det(bins, weights, n):
    """field.py:468-498 with det=True.  Returns (samples, inds)."""
    w = weights + 1e-5
    pdf = w / w.sum(-1, keepdim=True)
    cdf = torch.cat([torch.zeros_like(pdf[..., :1]), torch.cumsum(pdf, -1)], -1)
    u = torch.linspace(0.5 / n, 1.0 - 0.5 / n, n).expand(cdf.shape[0], n).contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    lo = torch.clamp(inds - 1, min=0)
    hi = torch.clamp(inds, max=cdf.shape[-1] - 1)
    c0, c1 = torch.gather(cdf, 1, lo), torch.gather(cdf, 1, hi)
    b0, b1 = torch.gather(bins, 1, lo), torch.gather(bins, 1, hi)
    den = c1 - c0
    den = torch.where(den < 1e-5, torch.ones_like(den), den)
    return b0 + (u - c0) / den * (b1 - b0), inds


def upsample_round(o, d, z, sdf, n_new, inv_s):
    """ZT:525-554 -- one SDF-guided importance round; returns (z_new[R,n_new], inds)."""
    pts = o[:, None, :] + d[:, None, :] * z[..., None]
    radius = torch.linalg.norm(pts, dim=-1)
    inside = (radius[:, :-1] < 1.0) | (radius[:, 1:] < 1.0)
    ps, ns = sdf[:, :-1], sdf[:, 1:]
    pz, nz = z[:, :-1], z[:, 1:]
    mid = (ps + ns) * 0.5
    cos = (ns - ps) / (nz - pz + 1e-5)
    prev = torch.cat([torch.zeros_like(cos[:, :1]), cos[:, :-1]], -1)
    cos = torch.minimum(prev, cos).clip(-1e3, 0.0) * inside
    dist = nz - pz
    pe, ne = mid - cos * dist * 0.5, mid + cos * dist * 0.5
    pc, nc = torch.sigmoid(pe * inv_s), torch.sigmoid(ne * inv_s)
    alpha = (pc - nc + 1e-5) / (pc + 1e-5)
    T = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1.0 - alpha + 1e-7], -1), -1)[:, :-1]
    return sample_pdf_det(z, alpha * T, n_new)


def sample_ray(sd, o, d, near, far, U0, U1, perturb=True, n_importance=64, up_steps=4, trace=None):
    """ZT:572-612.  `trace` (optional dict) receives the per-round intermediates."""
    z, z_out = coarse_samples(near, far, U0, U1, perturb=perturb)
    R = o.shape[0]
    with torch.no_grad():
        pts = o[:, None, :] + d[:, None, :] * z[..., None]
        sdf = sdf_forward(sd, pts.reshape(-1, 3))[:, 0].reshape(R, -1)
        inv_s_full = torch.exp(sd["deviation_network.variance"] * 10.0)
        for i in range(up_steps):
            inv_s = torch.clamp(inv_s_full, max=64.0 * 2 ** i)
            z_new, inds = upsample_round(o, d, z, sdf, n_importance // up_steps, inv_s)
            zc = torch.cat([z, z_new], -1)
            z_sorted, perm = torch.sort(zc, dim=-1)
            if trace is not None:
                trace[f"z_in_{i}"], trace[f"sdf_in_{i}"] = z, sdf
                trace[f"z_new_{i}"], trace[f"inds_{i}"], trace[f"perm_{i}"] = z_new, inds, perm
                trace[f"z_merged_{i}"] = z_sorted
            if i + 1 < up_steps:
                npts = o[:, None, :] + d[:, None, :] * z_new[..., None]
                new_sdf = sdf_forward(sd, npts.reshape(-1, 3))[:, 0].reshape(R, -1)
                sdf = torch.gather(torch.cat([sdf, new_sdf], -1), 1, perm)
            z = z_sorted
    return torch.cat([z, z_out], -1)


# ----------------------------------------------------------------------------- rendering
def composite(alpha, color):
    """ZT:773-775 -- w = alpha * exclusive_cumprod(1 - alpha + 1e-7); returns (w, sum w c)."""
    T = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1.0 - alpha + 1e-7], -1), -1)[:, :-1]
    w = alpha * T
    return w, (color * w[..., None]).sum(1)


# ----------------------------------------------------------------------------- occlusion-probe loss
def get_sphere_intersection(pts, dirs):
    """field.py:458-464."""
    dtx = (pts * dirs).sum(-1, keepdim=True)
    xtx = (pts ** 2).sum(-1, keepdim=True)
    dist = dtx ** 2 - xtx + 1
    return -dtx + torch.sqrt(dist + 1e-6)


def probe_weights(sd, z, origins, dirs, inv_s):
    """get_weights field.py:501-521 (sdf network + SingleVarianceNetwork of the stage-1 field)."""
    pts = z[..., None] * dirs[:, None, :] + origins[:, None, :]
    sdf = sdf_forward(sd, pts.reshape(-1, 3))[:, 0].reshape(z.shape)
    ps, ns = sdf[:, :-1], sdf[:, 1:]
    pz, nz = z[:, :-1], z[:, 1:]
    mid = (ps + ns) * 0.5
    cos = (ns - ps) / (nz - pz + 1e-5)
    surf = cos < 0
    cos = torch.clamp(cos, max=0)
    dist = nz - pz
    pe, ne = mid - cos * dist * 0.5, mid + cos * dist * 0.5
    pc, nc = torch.sigmoid(pe * inv_s), torch.sigmoid(ne * inv_s)
    alpha = (pc - nc + 1e-5) / (pc + 1e-5) * surf.float()
    T = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1.0 - alpha + 1e-7], -1), -1)[:, :-1]
    return alpha * T


def occ_probability(sd, pts, dirs, sn0=64, sn1=16):
    """get_intersection field.py:524-554 -> sum of the hit weights (occ_prob_gt of ZT:718-719).  pts must satisfy
    |p| < 0.999 (the caller's mask, ZT:704)."""
    with torch.no_grad():
        inv_s = torch.exp(sd["deviation_network.variance"] * 10.0)
        max_dist = get_sphere_intersection(pts, dirs)
        z = max_dist * torch.linspace(0, 1, sn0)[None, :]
        w = probe_weights(sd, z, pts, dirs, inv_s)
        z_new, _ = sample_pdf_det(z, w, sn1)
        w = probe_weights(sd, z_new, pts, dirs, inv_s)
    return w.sum(-1, keepdim=True)


def occ_loss(sd, info, points, sdf, grads, dirs, max_pn=2048, sdf_thresh=0.01, perm=None):
    """compute_occ_loss ZT:695-723.  `perm` = the torch.randperm draw used when more than max_pn samples qualify."""
    mask = (torch.norm(points, dim=-1) < 0.999) & ((grads * dirs).sum(-1) < 0) & (sdf.abs() < sdf_thresh)
    if mask.sum() > max_pn:
        idx = torch.nonzero(mask)[:, 0]
        idx = idx[perm[:max_pn]]
        mask = torch.zeros_like(mask)
        mask[idx] = True
    if mask.sum() == 0:
        return torch.zeros(1), mask
    gt = occ_probability(sd, points[mask], info["reflective"][mask].detach())
    return F.l1_loss(info["occ_prob"][mask], gt), mask


def render_core(sd, o, d, z, cos_anneal, step, is_nerf=True, freeze_inv_s_step=15000, exp_max=3.0,
                occ_loss_step=15000, occ_perm=None, occ_max_pn=2048):
    """ZT:725-820 (training outputs; the occlusion loss is active from occ_loss_step on)."""
    R, S = z.shape
    dists = z[:, 1:] - z[:, :-1]
    dists = torch.cat([dists, dists[:, -1:]], -1)
    zm = z + dists * 0.5
    pts = o[:, None, :] + d[:, None, :] * zm[..., None]
    inner = torch.norm(pts, dim=-1) <= 1.0
    outer = ~inner
    dirs = F.normalize(d[:, None, :].expand(R, S, 3), dim=-1)
    alpha = torch.zeros(R, S)
    color = torch.zeros(R, S, 3)
    out = {}
    if outer.any():
        p = pts[outer]
        nrm = torch.norm(p, dim=-1, keepdim=True)
        sig, rgb = nerfpp_forward(sd, torch.cat([p / nrm, 1.0 / nrm], -1), -dirs[outer])
        a_o = 1.0 - torch.exp(-F.softplus(sig[:, 0]) * dists[outer])
        c_o = linear_to_srgb(torch.exp(torch.clamp(rgb, max=5.0)))
        alpha = alpha.masked_scatter(outer, a_o)
        color = color.masked_scatter(outer[..., None].expand(-1, -1, 3), c_o)
    alpha_b, color_b = alpha, color
    if inner.any():
        p = pts[inner]
        dd = dirs[inner]
        y, grad = sdf_forward(sd, p, with_grad=True)
        sdf, feat = y[:, 0], y[:, 1:]
        inv_s = torch.exp(sd["deviation_network.variance"] * 10.0).clip(1e-6, 1e6)
        if freeze_inv_s_step is not None and step < freeze_inv_s_step:
            inv_s = inv_s.detach()
        tc = (dd * grad).sum(-1)
        ic = -(F.relu(-tc * 0.5 + 0.5) * (1.0 - cos_anneal) + F.relu(-tc) * cos_anneal)
        dist_i = dists[inner]
        en, ep = sdf + ic * dist_i * 0.5, sdf - ic * dist_i * 0.5
        pc, nc = torch.sigmoid(ep * inv_s), torch.sigmoid(en * inv_s)
        a_i = ((pc - nc + 1e-5) / (pc + 1e-5)).clip(0.0, 1.0)
        c_i, info = shading_forward(sd, p, grad, -dd, feat, exp_max=exp_max)
        alpha = alpha.masked_scatter(inner, a_i)
        color = color.masked_scatter(inner[..., None].expand(-1, -1, 3), c_i)
        out["gradient_error"] = (torch.linalg.norm(grad, dim=-1) - 1.0) ** 2
        out["std"] = torch.mean(1.0 / inv_s)
        out["transmission"] = info["transmission_weight"]
        out["metallic"] = info["metallic"]
        if step >= occ_loss_step:
            out["loss_occ"], out["occ_mask"] = occ_loss(sd, info, p, sdf, grad, dd, max_pn=occ_max_pn, perm=occ_perm)
        else:
            out["loss_occ"] = torch.zeros(1)
    else:
        out["loss_occ"] = torch.zeros(1)
        out["gradient_error"] = torch.zeros(1)
        out["std"] = torch.zeros(1)
    w, rgb = composite(alpha, color)
    _, rgb_b = composite(alpha_b, color_b)
    spec = linear_to_srgb(predictor(sd, "color_network.outer_light", ide(dirs[:, 0, :], torch.zeros(R, 1)),
                                    "exp", exp_max))
    acc = w.sum(-1)
    if is_nerf:
        rgb = rgb + (1.0 - acc[..., None])
    out.update({"ray_rgb": torch.clamp(rgb, 0.0, 1.0), "acc": acc, "color_bkgr": rgb_b, "color_spec": spec,
                "weights": w, "alpha": alpha, "sampled_color": color, "inner_mask": inner})
    return out


def render(sd, o, d, near, far, U0, U1, cos_anneal, step, is_nerf=True, perturb=True, trace=None):
    """NeROShapeRenderer.render (ZT:614-634)."""
    z = sample_ray(sd, o, d, near, far, U0, U1, perturb=perturb, trace=trace)
    out = render_core(sd, o, d, z, cos_anneal, step, is_nerf=is_nerf)
    out["z_vals"] = z
    return out


def charbonnier(pr, gt):
    """ZT:508-510."""
    return torch.sqrt(((gt - pr) ** 2).sum(-1) + 0.001)


def train_loss(out, gt, eikonal_weight=0.1, step=10000, occ_loss_step=15000):
    """trainer_zero.py:157-161 with the loss.py adapters of spherepot.yaml: rgb + eikonal (+ occ + outer_reg from
    occ_loss_step on: loss.py:97, :206-209 -- loss_occ * 1.0 and 0.5 * mse(color_bkgr, color_spec))."""
    loss = charbonnier(out["ray_rgb"], gt).mean() + (eikonal_weight * out["gradient_error"]).mean()
    if step >= occ_loss_step:
        loss = loss + out["loss_occ"].mean() + 0.5 * F.mse_loss(out["color_bkgr"], out["color_spec"])
    return loss


# ----------------------------------------------------------------------------- synthetic inputs
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

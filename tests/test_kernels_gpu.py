"""GPU parity tests of the individual kernels, all called through the C-ABI (nu_nerf_b200/_lib.py).

Bars: integer / index outputs bit-exact against oracle/sampling_oracle.c; floating point against a plain
torch fp32 reference with the tolerance written in each test.
"""
import ctypes

import numpy as np
import pytest
import torch

from conftest import np_ptr, sampling_tables

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ops():
    from nu_nerf_b200 import ops
    return ops


def _mk_planes(x, width, planes, rows=None):
    """fp32 [M,K] -> P with zero padding."""
    ops = _ops()
    M, K = x.shape
    rows = M if rows is None else rows
    p = ops.P(rows, width, planes, x.device, zero=True)
    ops.to_planes(x.contiguous(), p, M, K, False, 1.0, M, width)
    return p


# ----------------------------------------------------------------------------------------- dense layers
@pytest.mark.parametrize("planes", [1, 2])
@pytest.mark.parametrize("M,N,K,act", [(1000, 256, 256, 1), (4096 + 77, 256, 64, 2), (300, 16, 256, 0),
                                       (2048, 224, 256, 2), (5000, 128, 320, 1), (129, 256, 384, 0)])
def test_linear_matches_torch(planes, M, N, K, act):
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    x = torch.randn(M, K, device=DEV, generator=g)
    w = torch.randn(N, K, device=DEV, generator=g) / K ** 0.5
    b = torch.randn(N, device=DEV, generator=g) * 0.1
    A = _mk_planes(x, K, planes)
    B = _mk_planes(w, K, planes)
    out = ops.P(M, N, planes, DEV, zero=True)
    out32 = torch.zeros(M, N, device=DEV)
    ops.linear(A, B, M, N, K, bias=b, act=act, out=out, out_f32=out32)
    xr, wr = A.float(), B.float()          # the values the kernel actually saw
    ref = xr.double() @ wr.double().t() + b.double()
    if act == 1:
        ref = torch.relu(ref)
    elif act == 2:
        ref = torch.nn.functional.softplus(ref, beta=100)
    ref = ref.float()
    tol = 2e-5 if planes == 2 else 2e-5  # inputs are pre-rounded, so both modes see exact operands
    if planes == 2:
        # split mode drops the lo*lo term: relative 2^-16 per product
        tol = 1e-4
    err = (out32 - ref).abs().max().item()
    assert err < tol * max(1.0, ref.abs().max().item()), err
    back = out.float()
    rel = 2 ** -8 if planes == 1 else 2 ** -15
    assert (back - out32).abs().max().item() <= rel * out32.abs().max().item() + 1e-6


def test_linear_tc_equals_simt_and_epilogue_modes():
    ops = _ops()
    M, N, K = 777, 256, 256
    g = torch.Generator(device=DEV).manual_seed(5)
    x = torch.randn(M, K, device=DEV, generator=g)
    w = torch.randn(N, K, device=DEV, generator=g) / 16
    auxv = (torch.randn(M, N, device=DEV, generator=g) * 0.02).abs() * (torch.rand(M, N, device=DEV, generator=g) > 0.3)
    addv = torch.randn(M, N, device=DEV, generator=g)
    for planes in (1, 2):
        A, B = _mk_planes(x, K, planes), _mk_planes(w, K, planes)
        aux, add = _mk_planes(auxv, N, planes), _mk_planes(addv, N, planes)
        for mode in (1, 2):
            o1 = torch.zeros(M, N, device=DEV)
            ops.linear(A, B, M, N, K, aux=aux, aux_mode=mode, add=add, out_f32=o1, n_store=217, out_scale=0.5)
            ref = 0.5 * (A.float().double() @ B.float().double().t())
            av = aux.float().double()
            ref = ref * ((av > 0).double() if mode == 1 else (1 - torch.exp(-100 * av))) + add.float().double()
            ref[:, 217:] = 0
            assert (o1 - ref.float()).abs().max().item() < 2e-4 * ref.abs().max().item()
            old = ops.GEMM_IMPL
            try:
                ops.GEMM_IMPL = 1
                o2 = torch.zeros(M, N, device=DEV)
                ops.linear(A, B, M, N, K, aux=aux, aux_mode=mode, add=add, out_f32=o2, n_store=217, out_scale=0.5)
            finally:
                ops.GEMM_IMPL = old
            assert (o1 - o2).abs().max().item() < 2e-4 * ref.abs().max().item()


@pytest.mark.parametrize("planes", [1, 2])
@pytest.mark.parametrize("M,N,K", [(4096, 256, 256), (1000, 217, 256), (70000, 256, 64), (333, 3, 128), (2500, 256, 384)])
def test_linear_dw_matches_torch(planes, M, N, K):
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(M + N)
    z = torch.randn(M, N, device=DEV, generator=g)
    x = torch.randn(M, K, device=DEV, generator=g)
    Z = _mk_planes(z, ops.pad(N, 64), planes)
    X = _mk_planes(x, K, planes)
    dW = torch.zeros(ops.pad(N, 16), K, device=DEV)
    db = torch.zeros(ops.pad(N, 16), device=DEV)
    ops.linear_dw(Z, X, M, N, K, dW, db=db)          # bias gradient fused into the same launch
    ref = (Z.float()[:, :N].double().t() @ X.float().double()).float()
    scale = ref.abs().max().item()
    err = (dW[:N] - ref).abs().max().item()
    assert err < 2e-4 * scale, (err, scale)
    ref_b = Z.float()[:, :N].double().sum(0).float()
    assert (db[:N] - ref_b).abs().max().item() < 1e-3 * M ** 0.5
    assert db[N:].abs().max().item() == 0 if db.shape[0] > N else True
    bsum = torch.zeros(N, device=DEV)
    ops.colsum(Z, M, N, bsum)
    assert (bsum - ref_b).abs().max().item() < 1e-3 * M ** 0.5
    # accumulation semantics: a second call adds
    ops.linear_dw(Z, X, M, N, K, dW, db=db)
    assert (dW[:N] - 2 * ref).abs().max().item() < 4e-4 * scale
    assert (db[:N] - 2 * ref_b).abs().max().item() < 2e-3 * M ** 0.5


# ----------------------------------------------------------------------------------------- sampling
def _c_upsample(lib, o, d, z, sdf, n_new, inv_s, cap, u):
    R, n = z.shape
    z_new = np.zeros((R, n_new), np.float32)
    inds = np.zeros((R, n_new), np.int32)
    zm = np.zeros((R, n + n_new), np.float32)
    perm = np.zeros((R, n + n_new), np.int32)
    lib.oracle_upsample(np_ptr(o), np_ptr(d), np_ptr(z), np_ptr(sdf), R, n, n_new, ctypes.c_float(inv_s),
                        ctypes.c_float(cap), np_ptr(u), np_ptr(z_new), np_ptr(inds), np_ptr(zm), np_ptr(perm))
    return z_new, inds, zm, perm


@pytest.mark.parametrize("sphere,perturb", [(0, 1), (1, 0), (1, 1)])
def test_ray_setup_bit_exact(oracle_c, sphere, perturb):
    from nu_nerf_b200 import _lib
    from oracle import nunerf_oracle as orc
    R = 1031
    o, d = orc.synthetic_rays(R)
    U0, U1 = orc.synthetic_uniforms(R)
    tab = sampling_tables()
    near = np.full((R,), 0.8, np.float32)
    far = np.full((R,), 4.5, np.float32)
    z, zb = np.zeros((R, 64), np.float32), np.zeros((R, 32), np.float32)
    oracle_c.oracle_ray_setup(np_ptr(o.numpy()), np_ptr(d.numpy()), np_ptr(near), np_ptr(far), np_ptr(U0.numpy()),
                              np_ptr(U1.numpy()), np_ptr(tab.numpy()), R, sphere, perturb, np_ptr(z), np_ptr(zb))
    dn, df = torch.full((R,), 0.8, device=DEV), torch.full((R,), 4.5, device=DEV)
    dz, dzb = torch.zeros(R, 64, device=DEV), torch.zeros(R, 32, device=DEV)
    do, dd, dU0, dU1, dtab = o.to(DEV), d.to(DEV), U0.to(DEV), U1.to(DEV), tab.to(DEV)   # keep alive across the launch
    _lib.call("nunerf_ray_setup", do.data_ptr(), dd.data_ptr(), dn.data_ptr(), df.data_ptr(),
              dU0.data_ptr(), dU1.data_ptr(), dtab.data_ptr(), R, sphere, perturb, dz.data_ptr(),
              dzb.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(dz.cpu().numpy(), z) and np.array_equal(dzb.cpu().numpy(), zb)
    assert np.array_equal(dn.cpu().numpy(), near) and np.array_equal(df.cpu().numpy(), far)
    # and against the torch restatement of the reference arithmetic (ZT:580-594)
    nr, fr = torch.from_numpy(near)[:, None], torch.from_numpy(far)[:, None]
    zr, zbr = orc.coarse_samples(nr, fr, U0, U1, perturb=bool(perturb))
    assert (zr - torch.from_numpy(z)).abs().max().item() <= 5e-7
    assert ((zbr - torch.from_numpy(zb)).abs() / zbr.abs()).max().item() <= 3e-7


@pytest.mark.parametrize("n,n_new,cap", [(64, 16, 64.0), (80, 16, 128.0), (96, 16, 256.0), (112, 16, 512.0), (64, 32, 64.0), (97, 7, 1e9)])
def test_upsample_bit_exact_vs_c_oracle(oracle_c, n, n_new, cap):
    from nu_nerf_b200 import _lib
    from oracle import nunerf_oracle as orc
    R = 2053
    o, d = orc.synthetic_rays(R)
    g = torch.Generator().manual_seed(n)
    z = torch.sort(0.8 + 3.7 * torch.rand(R, n, generator=g), dim=-1)[0].contiguous()
    # a plausible sdf: distance to a sphere of radius 0.5 plus noise, sometimes with exact ties in z
    pts = o[:, None, :] + d[:, None, :] * z[..., None]
    sdf = (pts.norm(dim=-1) - 0.5 + 0.01 * torch.randn(R, n, generator=g)).contiguous()
    z[::7, 5] = z[::7, 4]
    u = torch.linspace(0.5 / n_new, 1 - 0.5 / n_new, n_new)
    inv_s = 20.0855
    ref = _c_upsample(oracle_c, o.numpy(), d.numpy(), z.numpy(), sdf.numpy(), n_new, inv_s, cap, u.numpy())
    dz_new = torch.zeros(R, n_new, device=DEV)
    dinds = torch.zeros(R, n_new, dtype=torch.int32, device=DEV)
    dzm = torch.zeros(R, n + n_new, device=DEV)
    dperm = torch.zeros(R, n + n_new, dtype=torch.int32, device=DEV)
    inv_dev = torch.tensor([inv_s], device=DEV)
    do, dd, dzz, dsdf, du = o.to(DEV), d.to(DEV), z.to(DEV), sdf.to(DEV), u.to(DEV)      # keep alive across the launch
    _lib.call("nunerf_upsample", do.data_ptr(), dd.data_ptr(), dzz.data_ptr(),
              dsdf.data_ptr(), R, n, n_new, inv_dev.data_ptr(), cap, du.data_ptr(), dz_new.data_ptr(),
              dinds.data_ptr(), dzm.data_ptr(), dperm.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(dinds.cpu().numpy(), ref[1]), "sample indices must be bit exact"
    assert np.array_equal(dz_new.cpu().numpy(), ref[0])
    assert np.array_equal(dzm.cpu().numpy(), ref[2])
    assert np.array_equal(dperm.cpu().numpy(), ref[3]), "merge permutation must be bit exact"
    # merged depths are sorted and the permutation is a permutation
    assert (dzm[:, 1:] >= dzm[:, :-1]).all()
    assert (torch.sort(dperm, dim=-1)[0].cpu() == torch.arange(n + n_new)[None]).all()
    # merge_sdf gathers by the permutation
    sdf_new = torch.randn(R, n_new, device=DEV)
    merged = torch.zeros(R, n + n_new, device=DEV)
    _lib.call("nunerf_merge_sdf", dsdf.data_ptr(), sdf_new.data_ptr(), dperm.data_ptr(), R, n, n_new,
              merged.data_ptr())
    assert torch.equal(merged, torch.gather(torch.cat([dsdf, sdf_new], -1), 1, dperm.long()))


# ----------------------------------------------------------------------------------------- geometry + compositing
def _geometry(o, d, z):
    from nu_nerf_b200 import _lib
    R, S = z.shape
    f = lambda *s: torch.zeros(*s, device=DEV)
    i32 = lambda *s: torch.zeros(*s, dtype=torch.int32, device=DEV)
    out = dict(dists=f(R, S), pts=f(R, S, 3), slot=i32(R, S), counts=i32(2), scratch=i32(2 * R),
               pts_in=f(R * S, 3), dists_in=f(R * S), dirs_in=f(R * S, 3), id_in=i32(R * S),
               pts_out=f(R * S, 3), dists_out=f(R * S), dirs_out=f(R * S, 3), id_out=i32(R * S), ray_map=i32(R, 10))
    _lib.call("nunerf_render_geometry", o.data_ptr(), d.data_ptr(), z.data_ptr(), R, S, *[out[k].data_ptr() for k in
              ("dists", "pts", "slot", "counts", "scratch", "pts_in", "dists_in", "dirs_in", "id_in", "pts_out",
               "dists_out", "dirs_out", "id_out", "ray_map")])
    return out


def _slot_from_ray_map(ray_map, R, S):
    """The per-sample slot array rebuilt (on the CPU) from the compact per-ray map: offsets + inner bit masks."""
    rm = ray_map.cpu().long()
    bits = ((rm[:, 2:, None] >> torch.arange(32)) & 1).reshape(R, -1)[:, :S].bool()
    in_idx = rm[:, 0:1] + torch.cumsum(bits.long(), 1) - bits.long()
    out_idx = rm[:, 1:2] + torch.cumsum((~bits).long(), 1) - (~bits).long()
    return torch.where(bits, in_idx, -1 - out_idx).int()


def test_render_geometry_and_compaction_order():
    from oracle import nunerf_oracle as orc
    R, S = 777, 160
    o, d = orc.synthetic_rays(R)
    g = torch.Generator().manual_seed(0)
    z = torch.sort(0.8 + 3.7 * torch.rand(R, S, generator=g), dim=-1)[0]
    do, dd, dzz = o.to(DEV), d.to(DEV), z.to(DEV).contiguous()
    out = _geometry(do, dd, dzz)
    torch.cuda.synchronize()
    dists = z[:, 1:] - z[:, :-1]
    dists = torch.cat([dists, dists[:, -1:]], -1)
    pts = o[:, None, :] + d[:, None, :] * (z + dists * 0.5)[..., None]
    inner = pts.norm(dim=-1) <= 1.0
    assert torch.equal(out["dists"].cpu(), dists)
    assert (out["pts"].cpu() - pts).abs().max().item() <= 5e-7
    gi = (out["slot"] >= 0).cpu()
    # the mask may differ only where ||p|| is within rounding of 1
    assert ((gi != inner) & ((pts.norm(dim=-1) - 1).abs() > 1e-6)).sum().item() == 0
    n_in = int(out["counts"][0])
    assert n_in == int(gi.sum()) and int(out["counts"][1]) == R * S - n_in
    # row-major mask order == boolean-mask indexing order of the reference (points[inner_mask])
    assert torch.equal(out["pts_in"][:n_in].cpu(), out["pts"].cpu()[gi])
    assert torch.equal(out["dists_in"][:n_in].cpu(), out["dists"].cpu()[gi])
    assert torch.equal(out["pts_out"][:R * S - n_in].cpu(), out["pts"].cpu()[~gi])
    assert torch.equal(out["id_in"][:n_in].cpu().long(), torch.nonzero(gi.reshape(-1))[:, 0])
    dn = torch.nn.functional.normalize(d, dim=-1)[:, None, :].expand(R, S, 3)
    assert (out["dirs_in"][:n_in].cpu() - dn[gi]).abs().max().item() <= 2e-7
    # the compact per-ray form of the map (what the engine passes to the compositing kernels) describes the same slots
    assert torch.equal(_slot_from_ray_map(out["ray_map"], R, S), out["slot"].cpu())
    # optional outputs: with dists / pts / slot / ids omitted the compact arrays and the map are unchanged
    from nu_nerf_b200 import _lib
    f = lambda *s: torch.zeros(*s, device=DEV)
    i32 = lambda *s: torch.zeros(*s, dtype=torch.int32, device=DEV)
    lean = dict(counts=i32(2), scratch=i32(2 * R), pts_in=f(R * S, 3), dists_in=f(R * S), dirs_in=f(R * S, 3),
                pts_out=f(R * S, 3), dists_out=f(R * S), dirs_out=f(R * S, 3), ray_map=i32(R, 10))
    _lib.call("nunerf_render_geometry", do.data_ptr(), dd.data_ptr(), dzz.data_ptr(), R, S, None, None, None,
              lean["counts"].data_ptr(), lean["scratch"].data_ptr(), lean["pts_in"].data_ptr(), lean["dists_in"].data_ptr(),
              lean["dirs_in"].data_ptr(), None, lean["pts_out"].data_ptr(), lean["dists_out"].data_ptr(),
              lean["dirs_out"].data_ptr(), None, lean["ray_map"].data_ptr())
    for k in ("counts", "pts_in", "dists_in", "dirs_in", "pts_out", "dists_out", "dirs_out", "ray_map"):
        assert torch.equal(lean[k], out[k]), k


@pytest.mark.parametrize("S", [160, 131, 200])
def test_composite_forward_and_backward(S):
    """Lane-per-sample kernels (slot form) and the staged kernels (per-ray map, S <= 160: 16-byte chunk staging through
    shared memory, 5 consecutive samples per lane) against torch autograd and the oracle."""
    from nu_nerf_b200 import _lib
    from oracle import nunerf_oracle as orc
    R = 515
    o, d = orc.synthetic_rays(R)
    g = torch.Generator().manual_seed(1)
    z = torch.sort(0.8 + 3.7 * torch.rand(R, S, generator=g), dim=-1)[0]
    do, dd, dzz = o.to(DEV), d.to(DEV), z.to(DEV).contiguous()
    geo = _geometry(do, dd, dzz)
    n_in, n_out = int(geo["counts"][0]), int(geo["counts"][1])
    inner = (geo["slot"] >= 0)
    a_in = (torch.rand(n_in, generator=g) ** 3).to(DEV).requires_grad_(True)
    c_in = torch.rand(n_in, 3, generator=g).to(DEV).requires_grad_(True)
    a_out = (torch.rand(n_out, generator=g) ** 3).to(DEV).requires_grad_(True)
    c_out = torch.rand(n_out, 3, generator=g).to(DEV).requires_grad_(True)
    a_in.data[::50] = 1.0  # opaque samples exercise the 1e-7 floor
    rgb, raw, acc, bk, w = (torch.zeros(R, 3, device=DEV), torch.zeros(R, 3, device=DEV), torch.zeros(R, device=DEV),
                            torch.zeros(R, 3, device=DEV), torch.zeros(R, S, device=DEV))
    _lib.call("nunerf_composite_fwd", a_in.data_ptr(), c_in.data_ptr(), a_out.data_ptr(), c_out.data_ptr(),
              geo["slot"].data_ptr(), R, S, 1, rgb.data_ptr(), raw.data_ptr(), acc.data_ptr(), bk.data_ptr(), w.data_ptr(),
              None)
    # same through the compact per-ray map (S <= 160: the staged kernel, another summation order)
    rgb2, raw2, acc2, bk2, w2 = (torch.zeros_like(rgb), torch.zeros_like(raw), torch.zeros_like(acc), torch.zeros_like(bk),
                                 torch.zeros_like(w))
    _lib.call("nunerf_composite_fwd", a_in.data_ptr(), c_in.data_ptr(), a_out.data_ptr(), c_out.data_ptr(),
              None, R, S, 1, rgb2.data_ptr(), raw2.data_ptr(), acc2.data_ptr(), bk2.data_ptr(), w2.data_ptr(),
              geo["ray_map"].data_ptr())
    for a_, b_ in ((rgb, rgb2), (raw, raw2), (acc, acc2), (bk, bk2), (w, w2)):
        assert (a_ - b_).abs().max().item() < 2e-6
    # training form: no dense weights output (NULL) -- same per-ray results, bit for bit
    rgb3, raw3, acc3, bk3 = torch.zeros_like(rgb), torch.zeros_like(raw), torch.zeros_like(acc), torch.zeros_like(bk)
    _lib.call("nunerf_composite_fwd", a_in.data_ptr(), c_in.data_ptr(), a_out.data_ptr(), c_out.data_ptr(),
              None, R, S, 1, rgb3.data_ptr(), raw3.data_ptr(), acc3.data_ptr(), bk3.data_ptr(), None,
              geo["ray_map"].data_ptr())
    for a_, b_ in ((rgb2, rgb3), (raw2, raw3), (acc2, acc3), (bk2, bk3)):
        assert torch.equal(a_, b_)
    # torch fp32 reference (ZT:773-788)
    alpha = torch.zeros(R, S, device=DEV).masked_scatter(inner, a_in).masked_scatter(~inner, a_out)
    color = torch.zeros(R, S, 3, device=DEV).masked_scatter(inner[..., None].expand(-1, -1, 3), c_in) \
        .masked_scatter((~inner)[..., None].expand(-1, -1, 3), c_out)
    w_ref, rgb_ref = orc.composite(alpha.cpu(), color.cpu())
    # (run the reference on the GPU tensors for autograd)
    T = torch.cumprod(torch.cat([torch.ones(R, 1, device=DEV), 1. - alpha + 1e-7], -1), -1)[:, :-1]
    wt = alpha * T
    acc_t = wt.sum(-1)
    rgb_t = torch.clamp((color * wt[..., None]).sum(1) + (1 - acc_t[..., None]), 0, 1)
    ab = alpha * (~inner)
    Tb = torch.cumprod(torch.cat([torch.ones(R, 1, device=DEV), 1. - ab + 1e-7], -1), -1)[:, :-1]
    bk_t = (color * (ab * Tb)[..., None]).sum(1)
    for rgb_k, acc_k, bk_k, w_k in ((rgb, acc, bk, w), (rgb2, acc2, bk2, w2)):
        assert (rgb_k - rgb_t).abs().max().item() < 2e-6 and (acc_k - acc_t).abs().max().item() < 2e-6
        assert (bk_k - bk_t).abs().max().item() < 2e-6 and (w_k - wt).abs().max().item() < 1e-6
        assert (w_k.cpu() - w_ref.detach()).abs().max().item() < 1e-6
    g_rgb = torch.randn(R, 3, generator=g).to(DEV)
    g_acc = torch.randn(R, generator=g).to(DEV)
    g_bk = torch.randn(R, 3, generator=g).to(DEV)
    ((rgb_t * g_rgb).sum() + (acc_t * g_acc).sum() + (bk_t * g_bk).sum()).backward()
    da_in, dc_in, da_out, dc_out = (torch.zeros_like(a_in), torch.zeros_like(c_in), torch.zeros_like(a_out),
                                    torch.zeros_like(c_out))
    _lib.call("nunerf_composite_bwd", a_in.data_ptr(), c_in.data_ptr(), a_out.data_ptr(), c_out.data_ptr(),
              geo["slot"].data_ptr(), R, S, 1, raw.data_ptr(), g_rgb.data_ptr(), g_acc.data_ptr(), g_bk.data_ptr(),
              da_in.data_ptr(), dc_in.data_ptr(), da_out.data_ptr(), dc_out.data_ptr(), None)
    for mine, ref in ((da_in, a_in.grad), (dc_in, c_in.grad), (da_out, a_out.grad), (dc_out, c_out.grad)):
        scale = ref.abs().max().item()
        assert (mine - ref).abs().max().item() < 1e-4 * scale, ((mine - ref).abs().max().item(), scale)
    m2 = [torch.zeros_like(x) for x in (da_in, dc_in, da_out, dc_out)]
    _lib.call("nunerf_composite_bwd", a_in.data_ptr(), c_in.data_ptr(), a_out.data_ptr(), c_out.data_ptr(),
              None, R, S, 1, raw.data_ptr(), g_rgb.data_ptr(), g_acc.data_ptr(), g_bk.data_ptr(),
              m2[0].data_ptr(), m2[1].data_ptr(), m2[2].data_ptr(), m2[3].data_ptr(), geo["ray_map"].data_ptr())
    for mine, ref in zip(m2, (a_in.grad, c_in.grad, a_out.grad, c_out.grad)):
        scale = ref.abs().max().item()
        assert (mine - ref).abs().max().item() < 1e-4 * scale, ((mine - ref).abs().max().item(), scale)


def test_staged_composite_with_arbitrary_inner_masks():
    """The staged kernels decode ANY inner/outer pattern (not only the single interval a sorted ray gives) and any run
    alignment: random masks, incl. rays without inner / without outer samples, against the slot-form kernels."""
    from nu_nerf_b200 import _lib
    R, S = 301, 160
    g = torch.Generator().manual_seed(5)
    bits = torch.rand(R, S, generator=g) < torch.rand(R, 1, generator=g)
    bits[0] = False
    bits[1] = True
    n_in_ray = bits.sum(1)
    in_off = torch.cumsum(n_in_ray, 0) - n_in_ray
    out_off = torch.arange(R) * S - in_off
    words = (bits.view(R, 5, 32).long() << torch.arange(32)).sum(-1)
    words = torch.where(words >= 2 ** 31, words - 2 ** 32, words)
    ray_map = torch.zeros(R, 10, dtype=torch.int32)
    ray_map[:, 0], ray_map[:, 1], ray_map[:, 2:7] = in_off.int(), out_off.int(), words.int()
    slot = _slot_from_ray_map(ray_map, R, S).to(DEV)
    ray_map = ray_map.to(DEV)
    n_in, n_out = int(bits.sum()), int((~bits).sum())
    a_in, c_in = (torch.rand(n_in, generator=g) ** 2).to(DEV), torch.rand(n_in, 3, generator=g).to(DEV)
    a_out, c_out = (torch.rand(n_out, generator=g) ** 2).to(DEV), torch.rand(n_out, 3, generator=g).to(DEV)
    outs = []
    for kw in ((slot.data_ptr(), None), (None, ray_map.data_ptr())):
        rgb, raw, acc, bk, w = (torch.zeros(R, 3, device=DEV), torch.zeros(R, 3, device=DEV), torch.zeros(R, device=DEV),
                                torch.zeros(R, 3, device=DEV), torch.zeros(R, S, device=DEV))
        _lib.call("nunerf_composite_fwd", a_in.data_ptr(), c_in.data_ptr(), a_out.data_ptr(), c_out.data_ptr(), kw[0], R, S,
                  1, rgb.data_ptr(), raw.data_ptr(), acc.data_ptr(), bk.data_ptr(), w.data_ptr(), kw[1])
        g_rgb, g_acc, g_bk = (torch.randn(R, 3, generator=torch.Generator().manual_seed(6)).to(DEV),
                              torch.randn(R, generator=torch.Generator().manual_seed(7)).to(DEV),
                              torch.randn(R, 3, generator=torch.Generator().manual_seed(8)).to(DEV))
        # gradients buffers with a canary border: nothing outside the lists may be written
        grads = [torch.full((n + 8,) + sh, 7.0, device=DEV) for n, sh in ((n_in, ()), (n_in, (3,)), (n_out, ()), (n_out, (3,)))]
        views = [t[4:-4] for t in grads]
        assert all(v.data_ptr() % 16 == 0 for v in views)
        _lib.call("nunerf_composite_bwd", a_in.data_ptr(), c_in.data_ptr(), a_out.data_ptr(), c_out.data_ptr(), kw[0], R, S,
                  1, raw.data_ptr(), g_rgb.data_ptr(), g_acc.data_ptr(), g_bk.data_ptr(), views[0].data_ptr(),
                  views[1].data_ptr(), views[2].data_ptr(), views[3].data_ptr(), kw[1])
        for t in grads:
            assert (t[:4] == 7.0).all() and (t[-4:] == 7.0).all()
        outs.append((rgb, raw, acc, bk, w, *[v.clone() for v in views]))
    for i, (x, y) in enumerate(zip(*outs)):
        tol = 2e-6 if i < 5 else 1e-4 * y.abs().max().item()
        assert (x - y).abs().max().item() < tol, i


# ----------------------------------------------------------------------------------------- tracing
def _uv_sphere(nu, nv, radius=0.6):
    th = torch.linspace(0, np.pi, nv + 1)
    ph = torch.linspace(0, 2 * np.pi, nu + 1)[:-1]
    v = [torch.tensor([[0.0, 0.0, radius]])]
    for i in range(1, nv):
        v.append(torch.stack([radius * torch.sin(th[i]) * torch.cos(ph), radius * torch.sin(th[i]) * torch.sin(ph),
                              radius * torch.cos(th[i]).expand(nu)], -1))
    v.append(torch.tensor([[0.0, 0.0, -radius]]))
    V = torch.cat(v, 0).float()
    F = []
    for j in range(nu):
        F.append([0, 1 + j, 1 + (j + 1) % nu])
    for i in range(1, nv - 1):
        a, b = 1 + (i - 1) * nu, 1 + i * nu
        for j in range(nu):
            j2 = (j + 1) % nu
            F.append([a + j, b + j, b + j2])
            F.append([a + j, b + j2, a + j2])
    last = V.shape[0] - 1
    a = 1 + (nv - 2) * nu
    for j in range(nu):
        F.append([a + j, last, a + (j + 1) % nu])
    return V, torch.tensor(F, dtype=torch.int32)


@pytest.mark.parametrize("nu,nv", [(24, 12), (224, 224)])
def test_bvh_trace_hit_ids_bit_exact(oracle_c, nu, nv):
    from nu_nerf_b200.tracer import TriangleBVH
    from oracle import nunerf_oracle as orc
    V, F = _uv_sphere(nu, nv)
    N = 4096 if nu > 100 else 2000
    o, d = orc.synthetic_rays(N)
    # half of the rays start inside the mesh (the second bounce of stage 2)
    o[::2] = 0.3 * o[::2] / 3.0
    bvh = TriangleBVH(V.to(DEV), F.to(DEV))
    hit, tri, t = bvh.trace(o.to(DEV), d.to(DEV), return_t=True)
    hb, tb, ttb = bvh.trace_brute(o.to(DEV), d.to(DEV))
    assert torch.equal(tri, tb) and torch.equal(hit, hb) and torch.equal(t, ttb)
    sub = slice(0, 512 if nu > 100 else N)
    tv = V[F.long()].reshape(-1, 9).contiguous().numpy()
    n = o[sub].shape[0]
    h = np.zeros(n, np.float32); ti = np.zeros(n, np.int32); tt = np.zeros(n, np.float32); uv = np.zeros((n, 2), np.float32)
    oracle_c.oracle_closest_hit(np_ptr(tv), tv.shape[0], np_ptr(o[sub].contiguous().numpy()),
                                np_ptr(d[sub].contiguous().numpy()), n, ctypes.c_float(1e16), np_ptr(h), np_ptr(ti),
                                np_ptr(tt), np_ptr(uv))
    assert np.array_equal(tri[sub].cpu().numpy(), ti), "triangle hit ids must be bit exact"
    assert np.array_equal(hit[sub].cpu().numpy(), h) and np.array_equal(t[sub].cpu().numpy(), tt)
    assert hit.sum().item() > 0.5 * N


@pytest.mark.parametrize("planes", [1, 2])
def test_linear_relu_bitmask_roundtrip(planes):
    """ReLU layers emit 1 bit per element; the backward GEMM masks with those bits instead of re-reading activations."""
    ops = _ops()
    M, N, K = 1000, 256, 128
    g = torch.Generator(device=DEV).manual_seed(9)
    x = torch.randn(M, K, device=DEV, generator=g)
    w = torch.randn(N, K, device=DEV, generator=g) / K ** 0.5
    A, B = _mk_planes(x, K, planes), _mk_planes(w, K, planes)
    out = ops.P(M, N, planes, DEV, zero=True)
    o32 = torch.zeros(M, N, device=DEV)
    mask = torch.zeros(M, 32, dtype=torch.uint8, device=DEV)
    ops.linear(A, B, M, N, K, act=1, out=out, out_f32=o32, mask_out=mask)
    bits = ((mask[:, :, None] >> torch.arange(8, device=DEV, dtype=torch.uint8)) & 1).reshape(M, 256).bool()
    assert torch.equal(bits, o32 > 0)
    assert (out.float() - o32).abs().max().item() <= (2 ** -8 if planes == 1 else 2 ** -15) * o32.abs().max().item() + 1e-6
    # backward-style GEMM: dX = (dZ W) masked by the bits
    dz = torch.randn(M, K, device=DEV, generator=g)
    Z = _mk_planes(dz, K, planes)
    WT = _mk_planes(w, K, planes)          # any [256, K] operand
    r1 = torch.zeros(M, N, device=DEV)
    ops.linear(Z, WT, M, N, K, mask_in=mask, out_f32=r1)
    ref = (Z.float().double() @ WT.float().double().t()).float() * bits
    assert (r1 - ref).abs().max().item() < 2e-4 * ref.abs().max().item()


# ------------------------------------------------------------------------------------------------ stage-2 segment kernels
def _srgb_to_linear_t(x):
    eps = torch.finfo(torch.float32).eps
    return torch.where(x <= 0.04045, 25.0 / 323.0 * x, ((200.0 * x + 11.0) / 211.0).clamp(min=eps) ** (12.0 / 5.0))


@pytest.mark.parametrize("N,S", [(1, 1), (37, 33), (500, 127), (1000, 255), (3, 256)])
def test_segment_compositing_forward_and_backward(N, S):
    """seg_composite_fwd/bwd_kernel against the reference's expression (ZT:1942-1951: cumprod weights, linear colour, sum,
    transmission) evaluated with torch in fp64 (forward 5e-6 -- 1 - alpha + 1e-7 itself rounds to 1 + 1.19e-7 in fp32, which
    a 255-term product turns into 1e-6 -- and gradients 2e-5 relative)."""
    from nu_nerf_b200.engine import SegCompositeFn
    g = torch.Generator().manual_seed(N * 1000 + S)
    alpha = torch.rand(N, S, generator=g) ** 3
    alpha[::7] = 0.0
    alpha[:, S // 2:] = alpha[:, S // 2:] * 0.05
    color = torch.rand(N, S, 3, generator=g)
    color[:, ::5] *= 0.04                                 # exercise the linear toe of the sRGB curve
    g_rgb, g_t = torch.randn(N, 3, generator=g), torch.randn(N, generator=g)
    a64, c64 = alpha.double().requires_grad_(True), color.double().requires_grad_(True)
    Tc = torch.cumprod(torch.cat([torch.ones(N, 1, dtype=torch.float64), 1.0 - a64 + 1e-7], -1), -1)
    rgb_ref = (_srgb_to_linear_t(c64) * (a64 * Tc[:, :-1])[..., None]).sum(1)
    t_ref = Tc[:, -1]
    ((rgb_ref * g_rgb.double()).sum() + (t_ref * g_t.double()).sum()).backward()
    a, c = alpha.to(DEV).requires_grad_(True), color.to(DEV).requires_grad_(True)
    rgb, t_end = SegCompositeFn.apply(a, c)
    ((rgb * g_rgb.to(DEV)).sum() + (t_end * g_t.to(DEV)).sum()).backward()
    assert (rgb.detach().cpu().double() - rgb_ref.detach()).abs().max().item() < 5e-6
    assert (t_end.detach().cpu().double() - t_ref.detach()).abs().max().item() < 5e-6
    for got, ref in ((a.grad, a64.grad), (c.grad, c64.grad)):
        scale = max(ref.abs().max().item(), 1e-12)
        assert (got.cpu().double() - ref).abs().max().item() < 2e-5 * scale + 1e-7


@pytest.mark.parametrize("R,n,n_new", [(1, 192, 64), (300, 192, 64), (77, 64, 16), (40, 256, 64)])
def test_alpha_importance_matches_the_reference_expression(R, n, n_new):
    """alpha_importance_kernel (upsample_nerf + cat_z_vals_nerf, ZT:1367-1397) against the reference's torch expression
    (cumprod / cumsum / searchsorted / sort).  The CDF inversion is ill-conditioned where the CDF is flat, and torch's scan
    order differs from the kernel's fixed one by ulps, so isolated samples may land in a neighbouring bin: 99.5 % of the
    merged depths within 1e-5 relative, every row sorted, every old sample present."""
    from nu_nerf_b200.engine import importance_merge
    g = torch.Generator().manual_seed(R + n)
    z = torch.linspace(0.1, 64.0, n).unsqueeze(0).expand(R, n).contiguous()
    alpha = (torch.rand(R, n, generator=g) ** 4) * 0.3
    T = torch.cumprod(torch.cat([torch.ones(R, 1), 1.0 - alpha + 1e-7], -1), -1)[:, :-1]
    w = (alpha * T)[:, :-1] + 1e-5
    pdf = w / w.sum(-1, keepdim=True)
    cdf = torch.cat([torch.zeros(R, 1), torch.cumsum(pdf, -1)], -1)
    u = torch.linspace(0.5 / n_new, 1.0 - 0.5 / n_new, n_new).expand(R, n_new).contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    lo, hi = torch.clamp(inds - 1, min=0), torch.clamp(inds, max=n - 1)
    c0, c1, b0, b1 = torch.gather(cdf, 1, lo), torch.gather(cdf, 1, hi), torch.gather(z, 1, lo), torch.gather(z, 1, hi)
    den = c1 - c0
    den = torch.where(den < 1e-5, torch.ones_like(den), den)
    ref = torch.sort(torch.cat([z, b0 + (u - c0) / den * (b1 - b0)], -1), dim=-1)[0]
    got = importance_merge(z.to(DEV), alpha.to(DEV), n_new).cpu()
    assert got.shape == (R, n + n_new)
    assert (got[:, 1:] >= got[:, :-1]).all()
    err = (got - ref).abs() / ref.abs().clamp_min(1.0)
    assert (err < 1e-5).float().mean().item() > 0.995, (err < 1e-5).float().mean().item()
    assert err.max().item() < 1.0                                   # never further than a couple of 0.33-wide bins
    # every original depth survives the merge
    for r in range(min(R, 8)):
        assert torch.isin(z[r], got[r]).all()

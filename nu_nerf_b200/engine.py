"""Stage-1 render engine: orchestration of the sm_100a kernels for NeROShapeRenderer.render
(reference network/renderer_zerothick.py:572-820, "ZT"; fields network/field.py).

`sample_ray`      -- ZT:572-612, no-grad: ray set-up kernel, SDF-only MLP passes, 4 fused up-sample/merge rounds.
`RenderCoreFn`    -- ZT:725-820 as ONE autograd node: the forward is an explicit launch sequence, the backward a
                     hand-derived reverse sequence (incl. the reverse-over-reverse of SDFNetwork.gradient); torch
                     only owns the parameters (weight-norm re-parametrisation happens outside, on tiny tensors).

precision: "split" = bf16 hi/lo planes, 3 tcgen05 MMAs per product, fp32 accumulate (meets the 1e-4 / 1e-3 parity
gates); "bf16" = single plane (the fast mode the benchmark reports; gradients within 2e-2).
"""
import ctypes as C
import math
import os

import torch
import torch.nn.functional as F

from . import _lib
from ._lib import call, ptr
from .ops import P, linear, linear_dw, colsum, f32_to_planes, chain, GEMM_IMPL, CHAIN_K0_MAX
from .weights import Dense, WeightBank

SQRT2 = math.sqrt(2.0)
F32 = torch.float32
# bf16 mode: evaluate predictor MLPs (and their dX chains) as ONE fused launch each (csrc/chain.cu) instead of one launch
# per layer.  NUNERF_FUSED_CHAINS=0 keeps the layer-by-layer path (A/B measurements; the split mode always uses it).
FUSED_CHAINS = os.environ.get("NUNERF_FUSED_CHAINS", "1") != "0"


def _fused(planes):
    return FUSED_CHAINS and planes == 1 and GEMM_IMPL == 0


def _f(*shape, dev):
    return torch.empty(*shape, dtype=F32, device=dev)


def _z(*shape, dev):
    return torch.zeros(*shape, dtype=F32, device=dev)


def sampling_tables(device):
    """linspace tables of ZT:580-590 / field.py:477, computed with torch exactly as the reference does."""
    t = torch.linspace(0.0, 1.0, 64)
    b = torch.linspace(1e-3, 1.0 - 1.0 / 33.0, 32)
    mids = 0.5 * (b[1:] + b[:-1])
    upper = torch.cat([mids, b[-1:]])
    lower = torch.cat([b[:1], mids])
    return torch.cat([t, lower, upper - lower, b]).contiguous().to(device)


# =============================================================================================== weights
def _wn(mod):
    """(v, g, bias) of a WNLinear / (weight, None, bias) of a PlainLinear."""
    if hasattr(mod, "weight_v"):
        return mod.weight_v, mod.weight_g, mod.bias
    return mod.weight, None, mod.bias


class SdfWeights:
    """Operands of the 9 SDF layers (field.py:64-131).  lin4 is pre-scaled by 1/sqrt(2) (the skip concat
    divides its input, field.py:143); lin8 is split into the feature rows (1..256) and the sdf row (0)."""

    def __init__(self, bank, net):
        lins = net.layers()
        self.L = []
        for l in range(8):
            v, g, b = _wn(lins[l])
            self.L.append(Dense(bank, v, g, b, scale=(1.0 / SQRT2 if l == 4 else 1.0)))
        v8, g8, b8 = _wn(lins[8])
        self.sdf_head = Dense(bank, v8, g8, b8, rows=(0, 1), need_t=False, row_f32=True)
        self.feat = Dense(bank, v8, g8, b8, rows=(1, 257), need_t=False)
        # backward sees lin8 as one layer with rows ordered [features (256), sdf (1)]
        self.cat8 = Dense(bank, v8, g8, None, row_rot=1, need_k=False, need_t=True, grad=False)

    @property
    def w_sdf(self):
        return self.sdf_head.row_f32


class PredW:
    """make_predictor (field.py:371-408): 4 dense layers at Sequential indices 0, 2, 4, 6."""

    def __init__(self, bank, seq, need_dx0=True):
        self.L = []
        for j, i in enumerate((0, 2, 4, 6)):
            v, g, b = _wn(seq[i])
            self.L.append(Dense(bank, v, g, b, need_t=(j > 0 or need_dx0)))
        self.n_out = self.L[3].N


class NerfW:
    """NeRFNetwork (field.py:212-263).  Layer 5 input is stored [h (256) | PE (84)] so the previous layer can write
    straight into it: its weight columns are rotated accordingly (reference order is [PE | h], field.py:276)."""

    def __init__(self, bank, net):
        self.pts = []
        for i in range(8):
            v, g, b = _wn(net.pts_linears[i])
            self.pts.append(Dense(bank, v, g, b, col_rot=(84 if i == 5 else 0)))     # WTk of layer 0: position gradient
        vf, _, bf = _wn(net.feature_linear)
        va, _, ba = _wn(net.alpha_linear)
        self.feat = Dense(bank, vf, None, bf, need_t=False)
        self.alpha = Dense(bank, va, None, ba, need_t=False)
        # backward operand [feature rows (256) ; alpha row (1)]^T, shared buffer filled by two descriptors
        self.cat8 = Dense(bank, vf, None, None, need_k=False, need_t=True, grad=False, t_cols=320)
        Dense(bank, va, None, None, need_k=False, need_t=False, grad=False, wtk_share=(self.cat8.WTk, 256))
        self.views = Dense(bank, *_wn(net.views_linears[0]))
        self.rgb = Dense(bank, *_wn(net.rgb_linear))


# =============================================================================================== SDF network
def sdf_infer_fused(w: SdfWeights, pts, timeline=None):
    """sdf(x) in ONE launch (csrc/chain.cu): PE + 9 layers with the activations resident on the SM (bf16 mode)."""
    M, dev = pts.shape[0], pts.device
    out = _f(M, dev=dev)
    a = _lib.SdfInferT()
    a.pts, a.M, a.sdf, a.ld_sdf = pts.data_ptr(), M, out.data_ptr(), 1
    a.timeline = ptr(timeline)
    lays = [w.L[l] for l in range(8)] + [w.sdf_head]
    for l, d in enumerate(lays):
        a.w[l], a.ldw[l], a.bias[l] = d.Wk.ptr, d.Wk.ld, d.b.data_ptr()
    call("nunerf_sdf_infer", C.byref(a))
    return out


def sdf_infer(w: SdfWeights, pts, planes, fused=None):
    """sdf(x) only (SDFNetwork.sdf, field.py:152) for the sampling passes: 8 hidden layers + the 1-row head.
    bf16 mode runs the fused chain kernel; the fp32-accurate split mode runs layer by layer (hi/lo planes)."""
    from .ops import GEMM_IMPL
    if fused is None:
        fused = planes == 1 and GEMM_IMPL == 0
    if fused:
        return sdf_infer_fused(w, pts.contiguous())
    M, dev = pts.shape[0], pts.device
    x0 = P(M, 64, planes, dev)
    call("nunerf_encode_pe", pts.data_ptr(), M, 3, 6, x0.ptr, x0.ld, x0.lo, 0, 0, 64)
    a = P(M, 256, planes, dev)
    b = P(M, 256, planes, dev)
    cat = P(M, 256, planes, dev)
    linear(x0, w.L[0].Wk, M, 256, 64, bias=w.L[0].b, act=2, out=a)
    linear(a, w.L[1].Wk, M, 256, 256, bias=w.L[1].b, act=2, out=b)
    linear(b, w.L[2].Wk, M, 256, 256, bias=w.L[2].b, act=2, out=a)
    linear(a, w.L[3].Wk, M, 224, 256, bias=w.L[3].b, act=2, out=cat, n_store=217)
    call("nunerf_encode_pe", pts.data_ptr(), M, 3, 6, cat.ptr, cat.ld, cat.lo, 217, 0, 39)
    linear(cat, w.L[4].Wk, M, 256, 256, bias=w.L[4].b, act=2, out=a)
    linear(a, w.L[5].Wk, M, 256, 256, bias=w.L[5].b, act=2, out=b)
    linear(b, w.L[6].Wk, M, 256, 256, bias=w.L[6].b, act=2, out=a)
    linear(a, w.L[7].Wk, M, 256, 256, bias=w.L[7].b, act=2, out=b)
    out = _f(M, 16, dev=dev)
    linear(b, w.sdf_head.Wk, M, 16, 256, bias=w.sdf_head.b, out_f32=out, n_store=1)
    return out[:, 0]


class SdfTape:
    pass


def sdf_forward(w: SdfWeights, pts, planes, xm: P):
    """Full forward + the adjoint ("gradient") pass.  Writes the 256 features into xm[:, 0:256]; returns the tape
    with sdf [M,16] (col 0), grad [M,3] and every activation needed by the backward."""
    M, dev = pts.shape[0], pts.device
    t = SdfTape()
    t.M, t.pts = M, pts
    t.x0 = P(M, 64, planes, dev)
    call("nunerf_encode_pe", pts.data_ptr(), M, 3, 6, t.x0.ptr, t.x0.ld, t.x0.lo, 0, 0, 64)
    A = [P(M, 256, planes, dev) for _ in range(8)]
    t.A = A
    Gs = [P(M, 256, planes, dev) for _ in range(8)]
    t.Gs = Gs
    if _fused(planes):
        return _sdf_forward_fused(w, t, xm)
    linear(t.x0, w.L[0].Wk, M, 256, 64, bias=w.L[0].b, act=2, out=A[0])
    linear(A[0], w.L[1].Wk, M, 256, 256, bias=w.L[1].b, act=2, out=A[1])
    linear(A[1], w.L[2].Wk, M, 256, 256, bias=w.L[2].b, act=2, out=A[2])
    linear(A[2], w.L[3].Wk, M, 224, 256, bias=w.L[3].b, act=2, out=A[3], n_store=217)   # A[3] = [a3 | PE]
    call("nunerf_encode_pe", pts.data_ptr(), M, 3, 6, A[3].ptr, A[3].ld, A[3].lo, 217, 0, 39)
    for l in range(4, 8):
        linear(A[l - 1], w.L[l].Wk, M, 256, 256, bias=w.L[l].b, act=2, out=A[l])
    linear(A[7], w.feat.Wk, M, 256, 256, bias=w.feat.b, out=xm, out_col=0)
    t.sdf = _f(M, 16, dev=dev)
    linear(A[7], w.sdf_head.Wk, M, 16, 256, bias=w.sdf_head.b, out_f32=t.sdf, n_store=1)
    # ---- adjoint pass: gs_l = d sdf / d z_l
    call("nunerf_rowvec_mask", w.w_sdf.data_ptr(), A[7].ptr, A[7].ld, A[7].lo, M, 256, Gs[7].ptr, Gs[7].ld, Gs[7].lo)
    for l in (7, 6, 5):
        linear(Gs[l], w.L[l].WTk, M, 256, 256, aux=A[l - 1], aux_mode=2, out=Gs[l - 1])
    u4 = _f(M, 256, dev=dev)
    linear(Gs[4], w.L[4].WTk, M, 256, 256, out_f32=u4)
    t.g_skip = _f(M, 39, dev=dev)
    call("nunerf_sdf_skip_split", u4.data_ptr(), A[3].ptr, A[3].ld, A[3].lo, M, Gs[3].ptr, Gs[3].ld, Gs[3].lo,
         t.g_skip.data_ptr())
    for l in (3, 2, 1):
        linear(Gs[l], w.L[l].WTk, M, 256, 256, aux=A[l - 1], aux_mode=2, out=Gs[l - 1])
    u0 = t.u0 = _f(M, 64, dev=dev)
    linear(Gs[0], w.L[0].WTk, M, 64, 256, out_f32=u0)
    t.grad = _f(M, 3, dev=dev)
    call("nunerf_sdf_grad_pe", pts.data_ptr(), u0.data_ptr(), 64, t.g_skip.data_ptr(), 39, M, t.grad.data_ptr())
    return t


def _sdf_forward_fused(w: SdfWeights, t: SdfTape, xm: P):
    """bf16 mode: the value pass is ONE launch (PE in-kernel, every activation written once for the backward), the adjoint
    pass two (the softplus' factors come from the stored activations in the epilogue: chain aux_mode 4)."""
    M, dev, pts, A, Gs = t.M, t.pts.device, t.pts, t.A, t.Gs
    t.sdf = _f(M, 16, dev=dev)
    lays = []
    for l in range(8):
        d = dict(W=w.L[l].Wk, N=256, K=64 if l == 0 else 256, bias=w.L[l].b, act=2, store=A[l], keep=1)
        if l == 3:
            d.update(N=224, n_real=217, cat_pe=1)          # A[3] = [a3 | PE]; lin4 carries the 1/sqrt(2)
        lays.append(d)
    lays.append(dict(W=w.sdf_head.Wk, N=16, K=256, bias=w.sdf_head.b, out32=t.sdf, n32=16))
    lays.append(dict(W=w.feat.Wk, N=256, K=256, bias=w.feat.b, store=xm))
    chain(None, M, 64, lays, pts=pts)
    # ---- adjoint pass: gs_l = d sdf / d z_l
    call("nunerf_rowvec_mask", w.w_sdf.data_ptr(), A[7].ptr, A[7].ld, A[7].lo, M, 256, Gs[7].ptr, Gs[7].ld, Gs[7].lo)
    u4 = _f(M, 256, dev=dev)
    chain(Gs[7], M, 256, [dict(W=w.L[l].WTk, N=256, K=256, aux_mode=4, aux1=A[l - 1], store=Gs[l - 1], keep=1)
                          for l in (7, 6, 5)] + [dict(W=w.L[4].WTk, N=256, K=256, out32=u4, n32=256)])
    t.g_skip = _f(M, 39, dev=dev)
    call("nunerf_sdf_skip_split", u4.data_ptr(), A[3].ptr, A[3].ld, A[3].lo, M, Gs[3].ptr, Gs[3].ld, Gs[3].lo,
         t.g_skip.data_ptr())
    u0 = t.u0 = _f(M, 64, dev=dev)
    chain(Gs[3], M, 256, [dict(W=w.L[l].WTk, N=256, K=256, aux_mode=4, aux1=A[l - 1], store=Gs[l - 1], keep=1)
                          for l in (3, 2, 1)] + [dict(W=w.L[0].WTk, N=64, K=256, out32=u0, n32=64)])
    t.grad = _f(M, 3, dev=dev)
    call("nunerf_sdf_grad_pe", pts.data_ptr(), u0.data_ptr(), 64, t.g_skip.data_ptr(), 39, M, t.grad.data_ptr())
    return t


def _sdf_position_grad(w: SdfWeights, t: SdfTape, g_pe0, g_pe4, d_grad):
    """d loss / d x of the SDF network's input point: the value network's two PE inputs (layer 0 and the skip concat of
    layer 4; g_pe4[:, 25:64] are the 39 skip columns, i.e. rows 217..255 of lin4's transposed operand read from row 192)
    plus the x-dependence of SDFNetwork.gradient itself (field.py:155-167): grad = J_pe(x)^T (u0 + g_skip), whose
    derivative with u held fixed is the PE Hessian term; the dependence of u on x is already inside g_pe0 / g_pe4
    because the reverse-over-reverse pass added E_l to every dZ_l."""
    M, dev = t.M, t.pts.device
    dx = _f(M, 3, dev=dev)
    call("nunerf_pe_bwd", t.pts.data_ptr(), 3, 6, g_pe0.data_ptr(), 64, g_pe4.data_ptr() + 25 * 4, 64, M, dx.data_ptr(), 0)
    call("nunerf_sdf_pe_hess", t.pts.data_ptr(), t.u0.data_ptr(), 64, t.g_skip.data_ptr(), 39, d_grad.data_ptr(), M,
         dx.data_ptr())
    return dx


def sdf_backward(w: SdfWeights, t: SdfTape, planes, dxm: P, d_sdf, d_grad, want_dx=False):
    """Backward of sdf_forward.  dxm[:, 0:256] holds d feat (planes); d_sdf [M], d_grad [M,3] fp32.
    Weight / bias gradients accumulate in the Dense objects (w.L[l].dW / .db, w.feat, w.sdf_head).
    want_dx: returns d loss / d pts [M,3] (the stage-2 path geometry is a function of the IoR network)."""
    M, dev = t.M, t.pts.device
    A, Gs = t.A, t.Gs
    dW = [w.L[l].dW for l in range(8)]
    db = [w.L[l].db for l in range(8)]
    if _fused(planes):
        return _sdf_backward_fused(w, t, dxm, d_sdf, d_grad, want_dx)
    g_pe0 = g_pe4 = None
    # ---- (a) reverse of the adjoint pass (forward-like chain on u~), produces E_l and the gs (x) u~ weight terms
    E = [P(M, 256, planes, dev) for _ in range(8)]
    ut = P(M, 64, planes, dev)
    ucat = P(M, 256, planes, dev)            # u~_4 = [u~ from layer 3 (217) | J d_grad (39)]
    call("nunerf_sdf_grad_pe_bwd", t.pts.data_ptr(), d_grad.data_ptr(), M, ut.ptr, ut.ld, ut.lo, 0, 64,
         ucat.ptr, ucat.ld, ucat.lo, 217, 39)
    gts = P(M, 256, planes, dev)
    u_a, u_b = P(M, 256, planes, dev), P(M, 256, planes, dev)
    u_in, K_in = ut, 64
    for l in range(8):
        N_l = 224 if l == 3 else 256
        linear(u_in, w.L[l].Wk, M, N_l, K_in, out=gts)
        linear_dw(Gs[l], u_in, M, 217 if l == 3 else 256, K_in, dW[l])
        if l == 3:
            u_next = ucat
        elif l == 7:
            u_next = u_a if u_in is not u_a else u_b
        else:
            u_next = u_a if u_in is not u_a else u_b
        n_real = 217 if l == 3 else 256
        call("nunerf_sdf_bwd2_ew", gts.ptr, gts.ld, gts.lo, A[l].ptr, A[l].ld, A[l].lo, Gs[l].ptr, Gs[l].ld, Gs[l].lo,
             M, 256, n_real, u_next.ptr, u_next.ld, u_next.lo, E[l].ptr, E[l].ld, E[l].lo)
        if l == 7:
            colsum(u_next, M, 256, w.sdf_head.dW[0])       # d w_sdf += sum_m gts_7 . s_7
        u_in, K_in = u_next, 256
    # ---- (b) backward of the value network; dZ8 = [d feat (256) | d sdf | 0...]
    f32_to_planes(d_sdf, dxm, M, 1, 64, col=256)
    linear_dw(dxm, A[7], M, 256, 256, w.feat.dW, db=w.feat.db)
    linear_dw(dxm, A[7], M, 1, 256, w.sdf_head.dW, z_col=256, db=w.sdf_head.db)
    dz = P(M, 256, planes, dev)
    dz2 = P(M, 256, planes, dev)
    linear(dxm, w.cat8.WTk, M, 256, 320, aux=A[7], aux_mode=2, add=E[7], out=dz)
    cur, other = dz, dz2
    for l in range(7, 0, -1):
        n_l = 217 if l == 3 else 256
        linear_dw(cur, A[l - 1], M, n_l, 256, dW[l], db=db[l])
        if l == 4:
            if want_dx:
                g_pe4 = _f(M, 64, dev=dev)
                linear(cur, w.L[4].WTk, M, 64, 256, b_row=192, out_f32=g_pe4)
            # input of lin4 is [a3 | PE]: only the first 217 columns carry on (PE has no parameters upstream)
            nxt = P(M, 256, planes, dev, zero=True)
            linear(cur, w.L[l].WTk, M, 224, 256, aux=A[3], aux_mode=2, add=E[3], out=nxt, n_store=217)
            cur = nxt
        else:
            linear(cur, w.L[l].WTk, M, 256, 256, aux=A[l - 1], aux_mode=2, add=E[l - 1], out=other)
            cur, other = other, cur
    linear_dw(cur, t.x0, M, 256, 64, dW[0], db=db[0])
    if want_dx:
        g_pe0 = _f(M, 64, dev=dev)
        linear(cur, w.L[0].WTk, M, 64, 256, out_f32=g_pe0)
        return _sdf_position_grad(w, t, g_pe0, g_pe4, d_grad)


def _sdf_backward_fused(w: SdfWeights, t: SdfTape, dxm: P, d_sdf, d_grad, want_dx=False):
    """bf16 mode: both reverse passes as fused chains (csrc/chain.cu aux_mode 5 / 6) -- the products gts_l and the
    intermediate dZ never round-trip through HBM between the GEMM and its element-wise glue; every u~_l / dZ_l is written
    once for its weight-gradient GEMM.  The two 217-wide layers around the skip concat keep the layer-by-layer path."""
    M, dev, planes = t.M, t.pts.device, 1
    A, Gs = t.A, t.Gs
    dW = [w.L[l].dW for l in range(8)]
    db = [w.L[l].db for l in range(8)]
    E = [P(M, 256, planes, dev) for _ in range(8)]
    U = [P(M, 256, planes, dev) for _ in range(8)]      # U[l] = u~_{l+1}: output of layer l of pass (a); U[3] = u~ cat
    ut = P(M, 64, planes, dev)
    call("nunerf_sdf_grad_pe_bwd", t.pts.data_ptr(), d_grad.data_ptr(), M, ut.ptr, ut.ld, ut.lo, 0, 64,
         U[3].ptr, U[3].ld, U[3].lo, 217, 39)
    # ---- (a) u~ chain: gts_l = u~_l W_l^T ; u~_{l+1} = gts_l . s_l ; E_l = gts_l . gs_l . 100 (1 - s_l)
    ew5 = lambda l, K: dict(W=w.L[l].Wk, N=256, K=K, aux_mode=5, aux1=A[l], aux2=Gs[l], e_out=E[l], store=U[l], keep=1)
    chain(ut, M, 64, [ew5(0, 64), ew5(1, 256), ew5(2, 256)])
    gts = P(M, 256, planes, dev)
    linear(U[2], w.L[3].Wk, M, 224, 256, out=gts)
    call("nunerf_sdf_bwd2_ew", gts.ptr, gts.ld, gts.lo, A[3].ptr, A[3].ld, A[3].lo, Gs[3].ptr, Gs[3].ld, Gs[3].lo,
         M, 256, 217, U[3].ptr, U[3].ld, U[3].lo, E[3].ptr, E[3].ld, E[3].lo)
    chain(U[3], M, 256, [ew5(l, 256) for l in (4, 5, 6, 7)])
    u_in, K_in = ut, 64
    for l in range(8):
        linear_dw(Gs[l], u_in, M, 217 if l == 3 else 256, K_in, dW[l])
        u_in, K_in = U[l], 256
    colsum(U[7], M, 256, w.sdf_head.dW[0])               # d w_sdf += sum_m gts_7 . s_7
    # ---- (b) backward of the value network; dZ8 = [d feat (256) | d sdf | 0...]
    f32_to_planes(d_sdf, dxm, M, 1, 64, col=256)
    linear_dw(dxm, A[7], M, 256, 256, w.feat.dW, db=w.feat.db)
    linear_dw(dxm, A[7], M, 1, 256, w.sdf_head.dW, z_col=256, db=w.sdf_head.db)
    DZ = [P(M, 256, planes, dev) for _ in range(8)]      # DZ[l] = d loss / d z_l
    DZ[3].t[:, 217:].zero_()                             # lin3 has 217 outputs: the tail columns are read as K padding
    ew6 = lambda l: dict(W=w.L[l].WTk, N=256, K=256, aux_mode=6, aux1=A[l - 1], aux2=E[l - 1], store=DZ[l - 1], keep=1)
    # dZ_7 = (dZ_8 [feat | sdf] W_8) . s_7 + E_7: the same glue as the layers below, as the chain's first layer on the
    # 320-column input when the kernel takes it (else one layer-kernel launch)
    head7 = []
    if CHAIN_K0_MAX >= 320:
        head7 = [dict(W=w.cat8.WTk, N=256, K=320, aux_mode=6, aux1=A[7], aux2=E[7], store=DZ[7], keep=1)]
    else:
        linear(dxm, w.cat8.WTk, M, 256, 320, aux=A[7], aux_mode=2, add=E[7], out=DZ[7])
    g_pe0 = g_pe4 = None
    if want_dx:
        g_pe0, g_pe4 = _f(M, 64, dev=dev), _f(M, 64, dev=dev)
    # (want_dx: the PE-input gradients are narrow heads on the resident dZ_4 / dZ_0 of the two chains)
    chain(dxm if head7 else DZ[7], M, 320 if head7 else 256, head7 + [ew6(7), ew6(6), ew6(5)] +
          ([dict(W=w.L[4].WTk, w_row=192, N=64, K=256, out32=g_pe4, n32=64)] if want_dx else []))
    # input of lin4 is [a3 | PE]: only the first 217 columns carry on (PE has no parameters upstream)
    linear(DZ[4], w.L[4].WTk, M, 224, 256, aux=A[3], aux_mode=2, add=E[3], out=DZ[3], n_store=217)
    chain(DZ[3], M, 256, [ew6(3), ew6(2), ew6(1)] +
          ([dict(W=w.L[0].WTk, N=64, K=256, out32=g_pe0, n32=64)] if want_dx else []))
    for l in range(7, 0, -1):
        linear_dw(DZ[l], A[l - 1], M, 217 if l == 3 else 256, 256, dW[l], db=db[l])
    linear_dw(DZ[0], t.x0, M, 256, 64, dW[0], db=db[0])
    if want_dx:
        return _sdf_position_grad(w, t, g_pe0, g_pe4, d_grad)


# =============================================================================================== predictors
class PredTape:
    pass


def pred_forward(w: PredW, x: P, M, K0, planes):
    dev = x.t.device
    t = PredTape()
    t.M, t.x, t.K0 = M, x, K0
    t.H = [P(M, 256, planes, dev) for _ in range(3)]
    t.Mk = [torch.empty(M, 32, dtype=torch.uint8, device=dev) for _ in range(3)]     # 1-bit ReLU masks
    t.head = _f(M, 16, dev=dev)
    t.fused = _fused(planes)
    t.perm = [0, 0, 0]           # masks written by the chain kernel are in its thread order (mask_perm)
    if t.fused:
        # one launch: hidden activations stay on the SM, each is written out once (dW needs it) with its ReLU mask
        def hidden(i, K):
            t.perm[i] = 1
            return dict(W=w.L[i].Wk, N=256, K=K, bias=w.L[i].b, act=1, mask_out=t.Mk[i], store=t.H[i], keep=1,
                        mask_perm=1)
        head = dict(W=w.L[3].Wk, N=16, K=256, bias=w.L[3].b, out32=t.head, n32=16)
        if K0 <= CHAIN_K0_MAX:
            chain(x, M, K0, [hidden(0, K0), hidden(1, 256), hidden(2, 256), head])
        else:
            linear(x, w.L[0].Wk, M, 256, K0, bias=w.L[0].b, act=1, out=t.H[0], mask_out=t.Mk[0])
            chain(t.H[0], M, 256, [hidden(1, 256), hidden(2, 256), head])
        return t
    linear(x, w.L[0].Wk, M, 256, K0, bias=w.L[0].b, act=1, out=t.H[0], mask_out=t.Mk[0])
    linear(t.H[0], w.L[1].Wk, M, 256, 256, bias=w.L[1].b, act=1, out=t.H[1], mask_out=t.Mk[1])
    linear(t.H[1], w.L[2].Wk, M, 256, 256, bias=w.L[2].b, act=1, out=t.H[2], mask_out=t.Mk[2])
    linear(t.H[2], w.L[3].Wk, M, 16, 256, bias=w.L[3].b, out_f32=t.head, n_store=w.n_out)
    return t


def pred_backward(w: PredW, t: PredTape, dz_head: P, planes, dx_planes: P = None, dx_add=False, dx_f32=None, dx_n=0,
                  dx_tail=None):
    """dz_head: planes [M,64] (n_out real columns).  Optionally produces dX of the first layer either as planes
    (accumulating when dx_add) or as fp32.  dx_tail = (b_row, fp32 [M,64]): additionally the 64 input columns starting
    at b_row (the position columns of the material predictors' [feature | p] input)."""
    M, dev = t.M, t.x.t.device
    gW = [w.L[i].dW for i in range(4)]
    gb = [w.L[i].db for i in range(4)]
    if t.fused:
        # dX chain in one launch: dZ_l = (dZ_{l+1} W_{l+1}) . relu'(z_l), each dZ_l written once for its dW GEMM
        d2, d1, d0 = P(M, 256, planes, dev), P(M, 256, planes, dev), P(M, 256, planes, dev)
        lays = [dict(W=w.L[3].WTk, N=256, K=64, mask_in=t.Mk[2], store=d2, keep=1, mask_perm=t.perm[2]),
                dict(W=w.L[2].WTk, N=256, K=256, mask_in=t.Mk[1], store=d1, keep=1, mask_perm=t.perm[1]),
                dict(W=w.L[1].WTk, N=256, K=256, mask_in=t.Mk[0], store=d0, keep=1, mask_perm=t.perm[0])]
        if dx_f32 is not None and dx_n % 16 == 0 and dx_n <= 256:
            lays.append(dict(W=w.L[0].WTk, N=dx_n, K=256, out32=dx_f32, n32=dx_n))
            dx_f32 = None
        if dx_tail is not None:
            # narrow head on the resident dZ_0 (keep = store = 0 leaves the activation in place)
            lays.append(dict(W=w.L[0].WTk, w_row=dx_tail[0], N=64, K=256, out32=dx_tail[1], n32=64))
            dx_tail = None
        chain(dz_head, M, 64, lays)
        linear_dw(dz_head, t.H[2], M, w.n_out, 256, gW[3], db=gb[3])
        linear_dw(d2, t.H[1], M, 256, 256, gW[2], db=gb[2])
        linear_dw(d1, t.H[0], M, 256, 256, gW[1], db=gb[1])
        linear_dw(d0, t.x, M, 256, t.K0, gW[0], db=gb[0])
        d2 = d0
    else:
        linear_dw(dz_head, t.H[2], M, w.n_out, 256, gW[3], db=gb[3])
        d2, d1 = P(M, 256, planes, dev), P(M, 256, planes, dev)
        linear(dz_head, w.L[3].WTk, M, 256, 64, mask_in=t.Mk[2], out=d2)
        linear_dw(d2, t.H[1], M, 256, 256, gW[2], db=gb[2])
        linear(d2, w.L[2].WTk, M, 256, 256, mask_in=t.Mk[1], out=d1)
        linear_dw(d1, t.H[0], M, 256, 256, gW[1], db=gb[1])
        linear(d1, w.L[1].WTk, M, 256, 256, mask_in=t.Mk[0], out=d2)
        linear_dw(d2, t.x, M, 256, t.K0, gW[0], db=gb[0])
    if dx_planes is not None:
        linear(d2, w.L[0].WTk, M, dx_n, 256, out=dx_planes, add=dx_planes if dx_add else None)
    if dx_f32 is not None:
        linear(d2, w.L[0].WTk, M, dx_n, 256, out_f32=dx_f32)
    if dx_tail is not None:
        linear(d2, w.L[0].WTk, M, 64, 256, b_row=dx_tail[0], out_f32=dx_tail[1])


# =============================================================================================== NeRF++
class NerfTape:
    pass


def nerf_forward(w: NerfW, pts, dirs, dists, planes):
    """compute_density_alpha (ZT:687-693) on the compact outer samples -> alpha [M], colour [M,3]."""
    M, dev = pts.shape[0], pts.device
    t = NerfTape()
    t.M, t.dists, t.pts = M, dists, pts
    pts4, views = _f(M, 4, dev=dev), _f(M, 3, dev=dev)
    t.pts4, t.views = pts4, views
    call("nunerf_nerf_prep", pts.data_ptr(), dirs.data_ptr(), M, pts4.data_ptr(), views.data_ptr())
    t.x0 = P(M, 128, planes, dev)
    call("nunerf_encode_pe", pts4.data_ptr(), M, 4, 10, t.x0.ptr, t.x0.ld, t.x0.lo, 0, 0, 128)
    H = [P(M, 256, planes, dev) for _ in range(8)]
    H[4] = P(M, 384, planes, dev)                 # [h4 | PE(84) | 0]
    t.H = H
    call("nunerf_encode_pe", pts4.data_ptr(), M, 4, 10, H[4].ptr, H[4].ld, H[4].lo, 256, 0, 128)
    t.Mk = [torch.empty(M, 32, dtype=torch.uint8, device=dev) for _ in range(8)]
    t.Mv = torch.empty(M, 16, dtype=torch.uint8, device=dev)
    t.xv = P(M, 320, planes, dev)                 # [feature | PE4(view) (27) | 0]
    t.sigma = _f(M, 16, dev=dev)
    t.perm = [0] * 8                              # masks written by the chain kernel are in its thread order
    if _fused(planes):
        def hid(i, K):
            t.perm[i] = 0 if i == 7 else 1        # Mk[7] is consumed by the layer kernel (K = 320 backward input)
            return dict(W=w.pts[i].Wk, N=256, K=K, bias=w.pts[i].b, act=1, mask_out=t.Mk[i], store=H[i], keep=1,
                        mask_perm=t.perm[i])
        chain(t.x0, M, 128, [hid(0, 128)] + [hid(i, 256) for i in range(1, 5)])
        linear(H[4], w.pts[5].Wk, M, 256, 384, bias=w.pts[5].b, act=1, out=H[5], mask_out=t.Mk[5])     # [h4 | PE]: K = 384
        chain(H[5], M, 256, [hid(6, 256), hid(7, 256),
                             dict(W=w.alpha.Wk, N=16, K=256, bias=w.alpha.b, out32=t.sigma, n32=16),
                             dict(W=w.feat.Wk, N=256, K=256, bias=w.feat.b, store=t.xv)])
    else:
        linear(t.x0, w.pts[0].Wk, M, 256, 128, bias=w.pts[0].b, act=1, out=H[0], mask_out=t.Mk[0])
        for i in range(1, 8):
            K = 384 if i == 5 else 256
            linear(H[i - 1], w.pts[i].Wk, M, 256, K, bias=w.pts[i].b, act=1, out=H[i], mask_out=t.Mk[i])
        linear(H[7], w.feat.Wk, M, 256, 256, bias=w.feat.b, out=t.xv)
        linear(H[7], w.alpha.Wk, M, 16, 256, bias=w.alpha.b, out_f32=t.sigma, n_store=1)
    call("nunerf_encode_pe", views.data_ptr(), M, 3, 4, t.xv.ptr, t.xv.ld, t.xv.lo, 256, 0, 64)
    t.hv = P(M, 128, planes, dev)
    linear(t.xv, w.views.Wk, M, 128, 320, bias=w.views.b, act=1, out=t.hv, mask_out=t.Mv)
    t.rgb = _f(M, 16, dev=dev)
    linear(t.hv, w.rgb.Wk, M, 16, 128, bias=w.rgb.b, out_f32=t.rgb, n_store=3)
    alpha, color = _f(M, dev=dev), _f(M, 3, dev=dev)
    call("nunerf_nerf_out_fwd", t.sigma.data_ptr(), 16, t.rgb.data_ptr(), 16, dists.data_ptr(), M, alpha.data_ptr(),
         color.data_ptr())
    return t, alpha, color


def nerf_backward(w: NerfW, t: NerfTape, d_alpha, d_color, planes, want_geo=False):
    """Reverse of nerf_forward.  want_geo: also returns (d_pts [M,3], d_dirs [M,3], d_dists [M]) -- the gradient with
    respect to the sample positions / ray directions / interval lengths (the stage-2 path geometry, ZT:1531-1539):
    dX GEMMs of the first trunk layer, the skip layer's PE rows and the view layer's PE rows, then the PE and
    inverted-sphere reverse kernels."""
    M, dev = t.M, t.dists.device
    H = t.H
    dz_rgb = P(M, 64, planes, dev, zero=True)
    dz8 = P(M, 320, planes, dev)
    dz8.t[:, 256:].zero_()                               # [d feature (256, written below) | d sigma | zero K padding]
    d_dists = _f(M, dev=dev) if want_geo else None
    if want_geo:
        call("nunerf_nerf_out_bwd_geo", t.sigma.data_ptr(), 16, t.rgb.data_ptr(), 16, t.dists.data_ptr(), M,
             d_alpha.data_ptr(), d_color.data_ptr(), dz8.ptr, dz8.ld, dz8.lo, 256, dz_rgb.ptr, dz_rgb.ld, dz_rgb.lo, 0,
             d_dists.data_ptr())
    else:
        call("nunerf_nerf_out_bwd", t.sigma.data_ptr(), 16, t.rgb.data_ptr(), 16, t.dists.data_ptr(), M,
             d_alpha.data_ptr(), d_color.data_ptr(), dz8.ptr, dz8.ld, dz8.lo, 256, dz_rgb.ptr, dz_rgb.ld, dz_rgb.lo, 0)
    linear_dw(dz_rgb, t.hv, M, 3, 128, w.rgb.dW, db=w.rgb.db)
    dzv = P(M, 128, planes, dev)
    linear(dz_rgb, w.rgb.WTk, M, 128, 64, mask_in=t.Mv, out=dzv)
    linear_dw(dzv, t.xv, M, 128, 320, w.views.dW, db=w.views.db)
    linear(dzv, w.views.WTk, M, 256, 128, out=dz8)                       # d feature -> dz8[:, :256]
    g_view = g_pe0 = g_pe5 = None
    if want_geo:
        g_view, g_pe0, g_pe5 = _f(M, 64, dev=dev), _f(M, 128, dev=dev), _f(M, 128, dev=dev)
        linear(dzv, w.views.WTk, M, 64, 128, b_row=256, out_f32=g_view)   # d PE4(view): rows 256.. of the view layer
    linear_dw(dz8, H[7], M, 256, 256, w.feat.dW, db=w.feat.db)
    linear_dw(dz8, H[7], M, 1, 256, w.alpha.dW, z_col=256, db=w.alpha.db)
    if _fused(planes) and t.perm[0]:
        # dZ_7 by the layer kernel (K = 320 input), then the whole dX chain of the 8 x 256 trunk in one launch
        DZ = [P(M, 256, planes, dev) for _ in range(8)]
        lays = []
        if CHAIN_K0_MAX >= 320:
            # dZ_7 from the 320-column [d feature | d sigma] rows as the chain's first layer
            lays.append(dict(W=w.cat8.WTk, N=256, K=320, mask_in=t.Mk[7], mask_perm=t.perm[7], store=DZ[7], keep=1))
        else:
            linear(dz8, w.cat8.WTk, M, 256, 320, mask_in=t.Mk[7], out=DZ[7])
        for i in range(7, 0, -1):
            lays.append(dict(W=w.pts[i].WTk, N=256, K=256, mask_in=t.Mk[i - 1], mask_perm=t.perm[i - 1], store=DZ[i - 1],
                             keep=1))
            if want_geo and i == 6:
                # narrow head on the resident dZ_5: the PE rows of the skip layer (its input is stored [h | PE])
                lays.append(dict(W=w.pts[5].WTk, w_row=256, N=128, K=256, out32=g_pe5, n32=128))
        if want_geo:
            lays.append(dict(W=w.pts[0].WTk, N=128, K=256, out32=g_pe0, n32=128))          # ... and on dZ_0
        if CHAIN_K0_MAX >= 320:
            chain(dz8, M, 320, lays)
        else:
            chain(DZ[7], M, 256, lays)
        for i in range(7, 0, -1):
            linear_dw(DZ[i], H[i - 1], M, 256, 384 if i == 5 else 256, w.pts[i].dW, db=w.pts[i].db)
        linear_dw(DZ[0], t.x0, M, 256, 128, w.pts[0].dW, db=w.pts[0].db)
    else:
        cur, other = P(M, 256, planes, dev), P(M, 256, planes, dev)
        linear(dz8, w.cat8.WTk, M, 256, 320, mask_in=t.Mk[7], out=cur)
        for i in range(7, 0, -1):
            K = 384 if i == 5 else 256
            linear_dw(cur, H[i - 1], M, 256, K, w.pts[i].dW, db=w.pts[i].db)
            if want_geo and i == 5:
                linear(cur, w.pts[5].WTk, M, 128, 256, b_row=256, out_f32=g_pe5)
            linear(cur, w.pts[i].WTk, M, 256, 256, mask_in=t.Mk[i - 1], out=other)
            cur, other = other, cur
        linear_dw(cur, t.x0, M, 256, 128, w.pts[0].dW, db=w.pts[0].db)
        if want_geo:
            linear(cur, w.pts[0].WTk, M, 128, 256, out_f32=g_pe0)
    if not want_geo:
        return None
    d_pts4, d_views = _f(M, 4, dev=dev), _f(M, 3, dev=dev)
    call("nunerf_pe_bwd", t.pts4.data_ptr(), 4, 10, g_pe0.data_ptr(), 128, g_pe5.data_ptr(), 128, M, d_pts4.data_ptr(), 0)
    call("nunerf_pe_bwd", t.views.data_ptr(), 3, 4, g_view.data_ptr(), 64, None, 0, M, d_views.data_ptr(), 0)
    d_pts, d_dirs = _f(M, 3, dev=dev), _f(M, 3, dev=dev)
    call("nunerf_nerf_prep_bwd", t.pts.data_ptr(), d_pts4.data_ptr(), d_views.data_ptr(), M, d_pts.data_ptr(),
         d_dirs.data_ptr())
    return d_pts, d_dirs, d_dists


# =============================================================================================== sampling
class Stage1Weights:
    """Persistent operands of the whole stage-1 field, refreshed from the module's parameters by ONE launch per step
    (WeightBank.prepare).  Built once per (module storage, precision)."""

    def __init__(self, net, planes, device):
        self.planes = planes
        self.bank = WeightBank(planes, device)
        self.sdf = SdfWeights(self.bank, net.sdf_network)
        self.nerf = NerfW(self.bank, net.outer_nerf)
        self.pred = {}
        for name, need_dx0 in (("metallic_predictor", True), ("roughness_predictor", True), ("albedo_predictor", True),
                               ("transmisstion_weight", True), ("outer_light", True), ("inner_light", True),
                               ("inner_weight", False), ("refrac_light", True)):
            self.pred[name] = PredW(self.bank, getattr(net.color_network, name), need_dx0=need_dx0)
        self.bank.finalize()
        self._variance = net.deviation_network.variance
        # encoding frequencies / refraction-light clamp of the shading network (AppShadingNetwork: 6 / 6 / light_exp_max;
        # AppShadingNetwork_SpecInner, field.py:1320-1330, :1373: 8 / 2 / -0.2)
        ccfg = getattr(net.color_network, "cfg", {})
        self.pos_freq, self.refrac_freq = int(ccfg.get("light_pos_freq", 6)), int(ccfg.get("refrac_freq", 6))
        if (self.pos_freq, self.refrac_freq) not in ((6, 6), (8, 2)):
            raise NotImplementedError("shading encode kernels: (light_pos_freq, refrac_freq) must be (6, 6) or (8, 2)")
        self.sphere_dir = bool(ccfg.get("sphere_direction", False))           # 144-wide outer-light input (field.py:594-597)
        self.exp_max_refrac = float(getattr(net.color_network.refrac_light, "exp_max", ccfg.get("light_exp_max", 3.0)))
        self.lut = net.color_network.FG_LUT.detach().float().contiguous()
        self.inv_s = torch.zeros(1, device=device)

    def refresh(self):
        self.bank.prepare()
        self.inv_s = torch.exp(self._variance.detach().float() * 10.0).reshape(1).contiguous()


_TABLES = {}


def _tables(device):
    key = str(device)
    if key not in _TABLES:
        _TABLES[key] = (sampling_tables(device),
                        {n: torch.linspace(0.5 / n, 1.0 - 0.5 / n, n).to(device) for n in (16, 32, 64)})
    return _TABLES[key]


@torch.no_grad()
def sample_ray(sw: SdfWeights, inv_s_dev, planes, rays_o, rays_d, near, far, perturb, U0=None, U1=None, sphere=False,
               n_importance=64, up_steps=4, trace=None):
    """ZT:572-612 -> z_vals [R,160].  U0 [R,1], U1 [R,32] are the uniform draws (torch.rand order of ZT:585,591)."""
    R, dev = rays_o.shape[0], rays_o.device
    tab, utabs = _tables(dev)
    o, d = rays_o.contiguous().float(), rays_d.contiguous().float()
    near, far = near.reshape(-1).contiguous().float().clone(), far.reshape(-1).contiguous().float().clone()
    z, z_bg = _f(R, 64, dev=dev), _f(R, 32, dev=dev)
    call("nunerf_ray_setup", o.data_ptr(), d.data_ptr(), near.data_ptr(), far.data_ptr(), ptr(U0), ptr(U1),
         tab.data_ptr(), R, int(sphere), int(perturb), z.data_ptr(), z_bg.data_ptr())
    n = 64
    pts = _f(R * n, 3, dev=dev)
    call("nunerf_points", o.data_ptr(), d.data_ptr(), z.data_ptr(), R, n, pts.data_ptr())
    sdf = sdf_infer(sw, pts, planes).reshape(R, n).contiguous()
    n_new = n_importance // up_steps
    u_tab = utabs[n_new]
    for i in range(up_steps):
        z_new, inds = _f(R, n_new, dev=dev), torch.empty(R, n_new, dtype=torch.int32, device=dev)
        z_m, perm = _f(R, n + n_new, dev=dev), torch.empty(R, n + n_new, dtype=torch.int32, device=dev)
        call("nunerf_upsample", o.data_ptr(), d.data_ptr(), z.data_ptr(), sdf.data_ptr(), R, n, n_new,
             inv_s_dev.data_ptr(), float(64 * 2 ** i), u_tab.data_ptr(), z_new.data_ptr(), inds.data_ptr(),
             z_m.data_ptr(), perm.data_ptr())
        if trace is not None:
            trace[f"z_in_{i}"], trace[f"sdf_in_{i}"] = z, sdf
            trace[f"z_new_{i}"], trace[f"inds_{i}"], trace[f"perm_{i}"], trace[f"z_merged_{i}"] = z_new, inds, perm, z_m
        if i + 1 < up_steps:
            npts = _f(R * n_new, 3, dev=dev)
            call("nunerf_points", o.data_ptr(), d.data_ptr(), z_new.data_ptr(), R, n_new, npts.data_ptr())
            sdf_new = sdf_infer(sw, npts, planes).reshape(R, n_new).contiguous()
            sdf_m = _f(R, n + n_new, dev=dev)
            call("nunerf_merge_sdf", sdf.data_ptr(), sdf_new.data_ptr(), perm.data_ptr(), R, n, n_new, sdf_m.data_ptr())
            sdf = sdf_m
        z, n = z_m, n + n_new
    return torch.cat([z, z_bg], -1).contiguous()


# =============================================================================================== render core
class CoreTape:
    pass


_MAT = ("metallic_predictor", "roughness_predictor", "albedo_predictor", "transmisstion_weight")


def inner_forward(w: Stage1Weights, t):
    """SDF network + sdf->alpha + shading on the compact inner samples t.pts_in / t.dists_in / t.dirs_in (ZT:759-769;
    stage 2: ZT:1887-1906 with the inner field).  Fills t.a_in, t.c_in, t.gerr and the tape entries."""
    planes, dev, M = w.planes, t.pts_in.device, t.n_in
    t.xm = P(M, 320, planes, dev)                                    # [feature (256) | p (3) | 0]
    f32_to_planes(t.pts_in, t.xm, M, 3, 64, col=256)
    t.sdf = sdf_forward(w.sdf, t.pts_in, planes, t.xm)
    t.a_in, t.gerr = _f(M, dev=dev), _f(M, dev=dev)
    sa = _lib.SdfAlphaT()
    sa.M, sa.cos_anneal, sa.inv_s_dev = M, t.cos_anneal, w.inv_s.data_ptr()
    sa.sdf, sa.ld_sdf, sa.grad, sa.dists, sa.dirs = t.sdf.sdf.data_ptr(), 16, t.sdf.grad.data_ptr(), \
        t.dists_in.data_ptr(), t.dirs_in.data_ptr()
    sa.alpha, sa.grad_err = t.a_in.data_ptr(), t.gerr.data_ptr()
    call("nunerf_sdf_alpha_fwd", C.byref(sa))
    shade_forward(w, t, t.sdf.grad)


def shade_forward(w: Stage1Weights, t, normals, no_refraction=False):
    """AppShadingNetwork.forward (field.py:684-741) on M = t.n_in points: t.xm = [feature | p] planes, t.pts_in,
    t.dirs_in (ray directions; view = -dirs), `normals` [M,3] (un-normalised: the SDF gradient in stage 1, the mesh
    normal for the stage-2 surface hits).  no_refraction: the stage-2 surface shader has no refraction-light term
    (field.py:966-967): its head is fed -inf so that exp(.) = 0 in the shared mixing kernel."""
    planes, dev, M = w.planes, t.pts_in.device, t.n_in
    # material predictors on [feature | p]
    t.mat = {k: pred_forward(w.pred[k], t.xm, M, 320, planes) for k in _MAT}
    # directions + encodings (the encode kernel writes whole 128-column rows, zero padded: no prior fill needed)
    ow = 192 if w.sphere_dir else 128            # sphere_direction: [IDE(u) | IDE(q(p, u)) | 0] (144 real columns)
    t.xo, t.xi = P(3 * M, ow, planes, dev), P(2 * M, 128, planes, dev)
    t.xw, t.xr = P(M, 128, planes, dev), P(M, 128, planes, dev)
    t.nov = _f(M, dev=dev)
    t.normals = normals
    se = _lib.ShadeEncodeT()
    se.M, se.pts, se.grad, se.dirs = M, t.pts_in.data_ptr(), normals.data_ptr(), t.dirs_in.data_ptr()
    se.rough_raw, se.ld_rough = t.mat["roughness_predictor"].head.data_ptr(), 16
    se.x_outer, se.ld_outer, se.lo_outer = t.xo.ptr, t.xo.ld, t.xo.lo
    se.x_inner, se.ld_inner, se.lo_inner = t.xi.ptr, t.xi.ld, t.xi.lo
    se.x_weight, se.ld_weight, se.lo_weight = t.xw.ptr, t.xw.ld, t.xw.lo
    se.x_refrac, se.ld_refrac, se.lo_refrac = t.xr.ptr, t.xr.ld, t.xr.lo
    se.nov = t.nov.data_ptr()
    se.pos_freq, se.refrac_freq, se.sphere_direction = w.pos_freq, w.refrac_freq, int(w.sphere_dir)
    t.refl = _f(M, 3, dev=dev)
    se.refl = t.refl.data_ptr()
    call("nunerf_shade_encode_fwd", C.byref(se))
    t.lo_ = pred_forward(w.pred["outer_light"], t.xo, 3 * M, ow, planes)
    t.li_ = pred_forward(w.pred["inner_light"], t.xi, 2 * M, 128, planes)
    t.lw_ = pred_forward(w.pred["inner_weight"], t.xw, M, 128, planes)
    if no_refraction:
        t.lr_ = PredTape()
        t.lr_.head = torch.full((M, 16), float("-inf"), device=dev)
    else:
        # (first-layer K padded to 64: 128 columns for the PE-6 inputs, 64 for the PE-2 ones of AppShadingNetwork_SpecInner)
        t.lr_ = pred_forward(w.pred["refrac_light"], t.xr, M, w.pred["refrac_light"].L[0].Kp, planes)
    t.c_in, t.trans, t.metallic, t.occ = _f(M, 3, dev=dev), _f(M, dev=dev), _f(M, dev=dev), _f(M, dev=dev)
    call("nunerf_shade_mix_fwd", C.byref(_mix_params(w, t)))


def sphere_exit_dir(p, u):
    """F.normalize(ps + u * get_sphere_intersection(ps, u)) with ps = offset_points_to_sphere(p) (field.py:447-465): where
    the ray (p, u) leaves the unit sphere.  [R,3] torch glue for the per-ray specular probe (NZ:806-808); the per-sample
    version lives in the encode kernels (pw::sphere_dir_fwd)."""
    pn = p.norm(dim=-1, keepdim=True)
    ps = torch.where(pn > 0.999, p / pn * 0.999, p)
    dtx = (ps * u).sum(-1, keepdim=True)
    xtx = (ps * ps).sum(-1, keepdim=True)
    t_ = -dtx + torch.sqrt(dtx * dtx - xtx + 1 + 1e-6)
    return F.normalize(ps + u * t_, dim=-1)


def core_forward(w: Stage1Weights, rays_o, rays_d, z_vals, cos_anneal, is_nerf, exp_max, want_weights=True, nz=False):
    """ZT:725-793 forward.  Returns the tape and the output tensors.  The dense per-sample compositing weights [R,S]
    are only needed by the validation outputs (depth / normal, ZT:657-693): a training step passes want_weights=False and
    the compositing kernel does not write them.
    nz: the stage-1 render_core of network/renderer.py (NZ:738-859) -- additionally t.loss_normal [R,1] = sum_j w_j
    max(grad_j . dir_j, 0) (NZ:766-780: a second pass of the compositing kernel with that scalar in the first colour
    channel), t.cand [R] = rays whose 65th mid-point lies in the unit sphere (NZ:798-802), and, with sphere_direction, the
    specular probe on [IDE(d, 0) | IDE(exit direction of (p_65, d), 0)] (NZ:805-809)."""
    if w.sphere_dir and not nz:
        raise NotImplementedError("stage-1 render_core with shader_config.sphere_direction: the per-ray specular probe over "
                                  "the rays whose 65th sample lies in the unit sphere (network/renderer.py:483-496) is not "
                                  "built; renderer_zerothick.py:780 itself feeds 72 columns to the 144-wide outer light")
    planes = w.planes
    R, S = z_vals.shape
    dev = z_vals.device
    t = CoreTape()
    t.R, t.S, t.is_nerf, t.cos_anneal, t.exp_max = R, S, int(is_nerf), float(cos_anneal), float(exp_max)
    o, d = rays_o.contiguous().float(), rays_d.contiguous().float()
    z = z_vals.contiguous().float()
    i32 = lambda *s: torch.empty(*s, dtype=torch.int32, device=dev)
    # compaction map in its per-ray form (offsets + inner bit masks, csrc/composite.cu RAY_MAP); the per-sample slot
    # array, the dense [R,S] dists / points and the flat sample ids are not needed by the step and are not written
    t.ray_map, counts, scratch = i32(R, 10), i32(2), i32(2 * R)
    cap = R * S
    pts_in, dists_in, dirs_in = _f(cap, 3, dev=dev), _f(cap, dev=dev), _f(cap, 3, dev=dev)
    pts_out, dists_out, dirs_out = _f(cap, 3, dev=dev), _f(cap, dev=dev), _f(cap, 3, dev=dev)
    call("nunerf_render_geometry", o.data_ptr(), d.data_ptr(), z.data_ptr(), R, S, None, None, None, counts.data_ptr(),
         scratch.data_ptr(), pts_in.data_ptr(), dists_in.data_ptr(), dirs_in.data_ptr(), None, pts_out.data_ptr(),
         dists_out.data_ptr(), dirs_out.data_ptr(), None, t.ray_map.data_ptr())
    n_in, n_out = (int(v) for v in counts.tolist())     # the one host sync of the step (output shapes need it)
    t.n_in, t.n_out = n_in, n_out
    t.pts_in, t.dists_in, t.dirs_in = pts_in[:n_in], dists_in[:n_in], dirs_in[:n_in]

    # ---- outer samples: NeRF++ (ZT:743-751)
    if n_out > 0:
        t.nerf, t.a_out, t.c_out = nerf_forward(w.nerf, pts_out[:n_out], dirs_out[:n_out], dists_out[:n_out], planes)
    else:
        t.nerf, t.a_out, t.c_out = None, _z(1, dev=dev), _z(1, 3, dev=dev)

    # ---- inner samples: SDF + shading (ZT:759-769)
    if t.n_in > 0:
        inner_forward(w, t)
    else:
        t.a_in, t.c_in, t.gerr = _z(1, dev=dev), _z(1, 3, dev=dev), None

    # ---- compositing (ZT:773-788)
    rgb, t.rgb_raw, acc, bkgr = _f(R, 3, dev=dev), _f(R, 3, dev=dev), _f(R, dev=dev), _f(R, 3, dev=dev)
    weights = _f(R, S, dev=dev) if want_weights else _f(0, S, dev=dev)
    call("nunerf_composite_fwd", t.a_in.data_ptr(), t.c_in.data_ptr(), t.a_out.data_ptr(), t.c_out.data_ptr(),
         None, R, S, t.is_nerf, rgb.data_ptr(), t.rgb_raw.data_ptr(), acc.data_ptr(), bkgr.data_ptr(),
         weights.data_ptr() if want_weights else None, t.ray_map.data_ptr())
    t.nz = nz
    if nz:
        if S < 66:
            raise ValueError("the non-zero-thickness stage-1 render_core reads sample 64 of every ray (NZ:798)")
        t.loss_normal = torch.zeros(R, 1, device=dev)
        if t.n_in > 0:
            t.nd_pos = (t.sdf.grad * t.dirs_in).sum(-1) > 0
            t.c_nd = torch.zeros(t.n_in, 3, device=dev)
            t.c_nd[:, 0] = torch.clamp((t.sdf.grad * t.dirs_in).sum(-1), min=0.0)
            t.c_zero = torch.zeros(max(t.n_out, 1), 3, device=dev)
            raw, scr3, scr1 = _f(R, 3, dev=dev), _f(R, 3, dev=dev), _f(R, dev=dev)
            call("nunerf_composite_fwd", t.a_in.data_ptr(), t.c_nd.data_ptr(), t.a_out.data_ptr(), t.c_zero.data_ptr(),
                 None, R, S, 0, scr3.data_ptr(), raw.data_ptr(), scr1.data_ptr(), _f(R, 3, dev=dev).data_ptr(), None,
                 t.ray_map.data_ptr())
            t.loss_normal = raw[:, :1].clone()
        zc = z[:, 64] + (z[:, 65] - z[:, 64]) * 0.5
        p_c = o + d * zc[:, None]
        t.cand = torch.norm(p_c, dim=-1) <= 1.0
    # ---- per-ray specular probe: outer_light(IDE(d, 0)) (ZT:780-781), activation applied by the caller
    t.dn = _norm_dirs(d)
    if w.sphere_dir:
        t.xs = P(R, 256, planes, dev, zero=True)      # [IDE(d, 0) (72) | IDE(q, 0) (72) | 0]: K = 192 of a 256-pitch row
        q = sphere_exit_dir(p_c, t.dn).contiguous()
        call("nunerf_ide_encode", t.dn.data_ptr(), R, 0.0, t.xs.ptr, t.xs.ld, t.xs.lo, 0)
        call("nunerf_ide_encode", q.data_ptr(), R, 0.0, t.xs.ptr, t.xs.ld, t.xs.lo, 72)
        t.ls_ = pred_forward(w.pred["outer_light"], t.xs, R, 192, planes)
    else:
        t.xs = P(R, 128, planes, dev)
        call("nunerf_ide_encode", t.dn.data_ptr(), R, 0.0, t.xs.ptr, t.xs.ld, t.xs.lo, 0)
        t.ls_ = pred_forward(w.pred["outer_light"], t.xs, R, 128, planes)
    return t, rgb, acc, bkgr, weights


def _norm_dirs(d):
    # F.normalize(dirs) of ZT:740 for the [R,3] ray directions (tiny, once per step)
    return (d / d.norm(dim=-1, keepdim=True).clamp_min(1e-12)).contiguous()


def _mix_params(w, t, d_color=None, d_trans=None, d_met=None, dz=None, d_rough=None, d_nov=None, d_occ=None):
    M = t.n_in
    mp = _lib.ShadeMixT()
    mp.M, mp.exp_max = M, t.exp_max
    mp.exp_max_refrac, mp.use_exp_max_refrac = w.exp_max_refrac, int(w.exp_max_refrac != t.exp_max)
    mp.metallic, mp.rough = t.mat["metallic_predictor"].head.data_ptr(), t.mat["roughness_predictor"].head.data_ptr()
    mp.albedo, mp.trans, mp.ld_mat = t.mat["albedo_predictor"].head.data_ptr(), \
        t.mat["transmisstion_weight"].head.data_ptr(), 16
    mp.outer, mp.ld_outer = t.lo_.head.data_ptr(), 16
    mp.inner, mp.ld_inner = t.li_.head.data_ptr(), 16
    mp.weight, mp.ld_weight = t.lw_.head.data_ptr(), 16
    mp.refrac, mp.ld_refrac = t.lr_.head.data_ptr(), 16
    mp.nov, mp.lut = t.nov.data_ptr(), w.lut.data_ptr()
    mp.color, mp.trans_out, mp.metallic_out, mp.occ_prob = t.c_in.data_ptr(), t.trans.data_ptr(), \
        t.metallic.data_ptr(), t.occ.data_ptr()
    if d_color is not None:
        mp.d_color, mp.d_trans_out, mp.d_metallic_out = d_color.data_ptr(), ptr(d_trans), ptr(d_met)
        mp.dz_metallic, mp.dz_albedo, mp.dz_trans = dz["metallic"].ptr, dz["albedo"].ptr, dz["trans"].ptr
        mp.dz_outer, mp.dz_inner, mp.dz_weight, mp.dz_refrac = dz["outer"].ptr, dz["inner"].ptr, dz["weight"].ptr, \
            dz["refrac"].ptr
        mp.ld_dz, mp.lo_dz = dz["metallic"].ld, dz["metallic"].lo
        mp.d_rough_raw, mp.d_nov = d_rough.data_ptr(), d_nov.data_ptr()
        mp.d_occ_prob = ptr(d_occ)
    return mp


def core_backward(w: Stage1Weights, t: CoreTape, d_rgb, d_acc, d_bkgr, d_gerr, d_trans, d_met, d_spec, want_inv_s,
                  d_occ=None, d_normal=None):
    """Reverse launch sequence of core_forward.  Returns {reference parameter name -> gradient of the EFFECTIVE
    weight / bias} (+ 'inv_s').  d_normal [R,1]: gradient of t.loss_normal (nz mode): the second compositing pass reversed
    -- its d alpha joins the main one, its d colour is d loss / d max(grad . dir, 0) and reaches the SDF gradient."""
    planes, dev = w.planes, t.ray_map.device
    R, S, M = t.R, t.S, t.n_in
    g = {}
    da_in, dc_in = _f(max(M, 1), dev=dev), _f(max(M, 1), 3, dev=dev)
    da_out, dc_out = _f(max(t.n_out, 1), dev=dev), _f(max(t.n_out, 1), 3, dev=dev)
    call("nunerf_composite_bwd", t.a_in.data_ptr(), t.c_in.data_ptr(), t.a_out.data_ptr(), t.c_out.data_ptr(),
         None, R, S, t.is_nerf, t.rgb_raw.data_ptr(), ptr(d_rgb), ptr(d_acc), ptr(d_bkgr),
         da_in.data_ptr(), dc_in.data_ptr(), da_out.data_ptr(), dc_out.data_ptr(), t.ray_map.data_ptr())
    d_grad_extra = None
    if d_normal is not None and M > 0:
        g3 = torch.zeros(R, 3, device=dev)
        g3[:, 0] = d_normal.reshape(-1)
        half = torch.full((R, 3), 0.5, device=dev)               # inside (0, 1): the clamp of the colour pass lets it through
        da2, dc2 = _f(max(M, 1), dev=dev), _f(max(M, 1), 3, dev=dev)
        da3, dc3 = _f(max(t.n_out, 1), dev=dev), _f(max(t.n_out, 1), 3, dev=dev)
        call("nunerf_composite_bwd", t.a_in.data_ptr(), t.c_nd.data_ptr(), t.a_out.data_ptr(), t.c_zero.data_ptr(),
             None, R, S, 0, half.data_ptr(), g3.data_ptr(), None, None,
             da2.data_ptr(), dc2.data_ptr(), da3.data_ptr(), dc3.data_ptr(), t.ray_map.data_ptr())
        da_in += da2
        if t.n_out > 0:
            da_out += da3
        d_grad_extra = (dc2[:M, 0] * t.nd_pos)[:, None] * t.dirs_in
    if t.n_out > 0:
        nerf_backward(w.nerf, t.nerf, da_out, dc_out, planes)

    # ---- specular probe (outer_light on the ray directions)
    if d_spec is not None:
        dzs = P(R, 64, planes, dev, zero=True)
        f32_to_planes(d_spec.contiguous(), dzs, R, 3, 64)
        pred_backward(w.pred["outer_light"], t.ls_, dzs, planes)
    if M == 0:
        return g
    inner_backward(w, t, da_in, dc_in, d_gerr, d_trans, d_met, want_inv_s, d_occ, g, d_grad_extra=d_grad_extra)
    return g


def inner_backward(w: Stage1Weights, t, da_in, dc_in, d_gerr, d_trans, d_met, want_inv_s, d_occ=None, g=None,
                   surface=False, want_geo=False, d_nov_ext=None, d_grad_extra=None):
    """Reverse of inner_forward on the M = t.n_in compact samples: shading mix, the light / material predictors,
    sdf -> alpha, the direction encodings and the SDF network (value pass + reverse-over-reverse of its gradient).
    surface=True is the reverse of the stage-2 surface shading (shade_forward with the MESH normal, no refraction
    light, no sdf -> alpha): the normal is a constant of the FIELD there, so only the feature path reaches the SDF
    network.  want_geo (stage 2): g["d_pts"], g["d_dirs"] [M,3] and g["d_dists"] [M] (inner) / g["d_normals"] [M,3]
    (surface) -- the gradient with respect to the sample geometry, which the IoR network moves."""
    planes, dev, M = w.planes, t.pts_in.device, t.n_in
    g = {} if g is None else g
    # ---- shading mix
    # (shade_mix_bwd writes every row of these dZ operands whole, zero padding included: no prior fill)
    dz = {k: P(M, 64, planes, dev) for k in ("metallic", "albedo", "trans", "weight", "refrac")}
    dz["outer"], dz["inner"] = P(3 * M, 64, planes, dev), P(2 * M, 64, planes, dev)
    d_rough, d_nov = _f(M, dev=dev), _f(M, dev=dev)
    call("nunerf_shade_mix_bwd", C.byref(_mix_params(w, t, dc_in, d_trans, d_met, dz, d_rough, d_nov, d_occ)))
    if d_nov_ext is not None:
        d_nov += d_nov_ext
    ow = 192 if w.sphere_dir else 128
    dxo, dxi = _f(3 * M, ow, dev=dev), _f(2 * M, 128, dev=dev)
    pred_backward(w.pred["outer_light"], t.lo_, dz["outer"], planes, dx_f32=dxo, dx_n=ow)
    pred_backward(w.pred["inner_light"], t.li_, dz["inner"], planes, dx_f32=dxi, dx_n=128)
    pred_backward(w.pred["inner_weight"], t.lw_, dz["weight"], planes)
    dxr = None
    if not surface:
        if want_geo:
            kr = w.pred["refrac_light"].L[0].Kp
            dxr = _f(M, kr, dev=dev)
            pred_backward(w.pred["refrac_light"], t.lr_, dz["refrac"], planes, dx_f32=dxr, dx_n=kr)
        else:
            pred_backward(w.pred["refrac_light"], t.lr_, dz["refrac"], planes)
    # ---- sdf -> alpha
    d_sdf, d_grad = _z(M, dev=dev), _z(M, 3, dev=dev)
    d_pts = _z(M, 3, dev=dev) if want_geo else None
    d_dirs = _z(M, 3, dev=dev) if want_geo else None
    d_dists = _z(M, dev=dev) if want_geo else None
    if not surface:
        d_inv = _z(1, dev=dev) if want_inv_s else None
        sa = _lib.SdfAlphaT()
        sa.M, sa.cos_anneal, sa.inv_s_dev = M, t.cos_anneal, w.inv_s.data_ptr()
        sa.sdf, sa.ld_sdf, sa.grad, sa.dists, sa.dirs = t.sdf.sdf.data_ptr(), 16, t.sdf.grad.data_ptr(), \
            t.dists_in.data_ptr(), t.dirs_in.data_ptr()
        sa.d_alpha, sa.d_grad_err = da_in.data_ptr(), ptr(d_gerr)
        sa.d_sdf, sa.d_grad, sa.d_inv_s = d_sdf.data_ptr(), d_grad.data_ptr(), ptr(d_inv)
        sa.d_dists, sa.d_dirs = ptr(d_dists), ptr(d_dirs)
        call("nunerf_sdf_alpha_bwd", C.byref(sa))
        if want_inv_s:
            g["inv_s"] = d_inv
    if d_grad_extra is not None:            # gradient that reaches the SDF gradient directly (loss_normal, NZ:766-780);
        d_grad += d_grad_extra              # after sdf_alpha_bwd, which WRITES d_grad
    # ---- directions / encodings (adds into d_grad and d_rough; with want_geo also into d_pts / d_dirs)
    se = _lib.ShadeEncodeT()
    se.M, se.pts, se.grad, se.dirs = M, t.pts_in.data_ptr(), t.normals.data_ptr(), t.dirs_in.data_ptr()
    se.rough_raw, se.ld_rough = t.mat["roughness_predictor"].head.data_ptr(), 16
    se.d_x_outer, se.ld_dxo, se.d_x_inner, se.ld_dxi = dxo.data_ptr(), dxo.stride(0), dxi.data_ptr(), 128
    se.d_nov, se.d_grad, se.d_rough_raw, se.ld_drough = d_nov.data_ptr(), d_grad.data_ptr(), d_rough.data_ptr(), 1
    se.pos_freq, se.refrac_freq, se.sphere_direction = w.pos_freq, w.refrac_freq, int(w.sphere_dir)
    if want_geo:
        se.d_x_refrac, se.ld_dxr = ptr(dxr), (dxr.stride(0) if dxr is not None else 0)
        se.d_pts, se.d_dirs = d_pts.data_ptr(), d_dirs.data_ptr()
    call("nunerf_shade_encode_bwd", C.byref(se))
    if surface:
        if want_geo:
            g["d_normals"] = d_grad
        d_grad = _z(M, 3, dev=dev)                                  # the mesh normal is not a function of the field
    dz["rough"] = P(M, 64, planes, dev)
    f32_to_planes(d_rough, dz["rough"], M, 1, 64)
    # ---- material predictors, d feature accumulated in dxm[:, :256]
    dxm = P(M, 320, planes, dev)
    first = True
    for name, key in (("metallic_predictor", "metallic"), ("roughness_predictor", "rough"),
                      ("albedo_predictor", "albedo"), ("transmisstion_weight", "trans")):
        tail = (256, _f(M, 64, dev=dev)) if want_geo else None       # d of the p columns of [feature | p]
        pred_backward(w.pred[name], t.mat[name], dz[key], planes, dx_planes=dxm, dx_add=not first, dx_n=256, dx_tail=tail)
        if want_geo:
            d_pts += tail[1][:, :3]
        first = False
    # ---- SDF network
    dx = sdf_backward(w.sdf, t.sdf, planes, dxm, d_sdf, d_grad, want_dx=want_geo)
    if want_geo:
        g["d_pts"], g["d_dirs"], g["d_dists"] = d_pts + dx, d_dirs, d_dists
    return g


# =============================================================================================== stage 2
class IorWeights:
    """IoRNetwork (field.py:1046-1065): PE-6 -> 256 relu -> 256 relu -> 256 -> 1, sigmoid (Sequential 0, 2, 4, 5)."""

    def __init__(self, net, planes, device):
        self.planes = planes
        self.bank = WeightBank(planes, device)
        self.L = [Dense(self.bank, *_wn(net.module0[i]), need_t=False, grad=False) for i in (0, 2, 4, 5)]
        self.bank.finalize()

    def refresh(self):
        self.bank.prepare()


@torch.no_grad()
def ior_forward(w: IorWeights, pts):
    """IoRNetwork.forward on [M,3] points -> [M] (after the sigmoid)."""
    M, dev, planes = pts.shape[0], pts.device, w.planes
    x0 = P(M, 64, planes, dev)
    call("nunerf_encode_pe", pts.data_ptr(), M, 3, 6, x0.ptr, x0.ld, x0.lo, 0, 0, 64)
    a, b = P(M, 256, planes, dev), P(M, 256, planes, dev)
    linear(x0, w.L[0].Wk, M, 256, 64, bias=w.L[0].b, act=1, out=a)
    linear(a, w.L[1].Wk, M, 256, 256, bias=w.L[1].b, act=1, out=b)
    linear(b, w.L[2].Wk, M, 256, 256, bias=w.L[2].b, act=0, out=a)
    out = _f(M, 16, dev=dev)
    linear(a, w.L[3].Wk, M, 16, 256, bias=w.L[3].b, out_f32=out, n_store=1)
    return torch.sigmoid(out[:, 0])


def segment_points(start, delta, z):
    """[R, n, 3] points start + delta * z (ZT:1727-1731, :1760, :1799)."""
    R, n = z.shape
    pts = _f(R, n, 3, dev=z.device)
    if R > 0:
        call("nunerf_points", start.contiguous().data_ptr(), delta.contiguous().data_ptr(), z.contiguous().data_ptr(),
             R, n, pts.data_ptr())
    return pts


@torch.no_grad()
def upsample_rounds(w: Stage1Weights, o, d, z, sdf, n_new, rounds):
    """`rounds` SDF-guided importance rounds with `n_new` samples each on the field `w` (ZT:1748-1758: inv_s capped at
    64 * 2^i, new SDF values queried at o + d * z_new, no query after the last round).  Returns the merged z."""
    R, n = z.shape
    dev = z.device
    _, utabs = _tables(dev)
    u_tab = utabs[n_new]
    for i in range(rounds):
        z_new, inds = _f(R, n_new, dev=dev), torch.empty(R, n_new, dtype=torch.int32, device=dev)
        z_m, perm = _f(R, n + n_new, dev=dev), torch.empty(R, n + n_new, dtype=torch.int32, device=dev)
        call("nunerf_upsample", o.data_ptr(), d.data_ptr(), z.data_ptr(), sdf.data_ptr(), R, n, n_new,
             w.inv_s.data_ptr(), float(64 * 2 ** i), u_tab.data_ptr(), z_new.data_ptr(), inds.data_ptr(),
             z_m.data_ptr(), perm.data_ptr())
        if i + 1 < rounds:
            npts = segment_points(o, d, z_new)
            sdf_new = sdf_infer(w.sdf, npts.reshape(-1, 3), w.planes).reshape(R, n_new).contiguous()
            sdf_m = _f(R, n + n_new, dev=dev)
            call("nunerf_merge_sdf", sdf.data_ptr(), sdf_new.data_ptr(), perm.data_ptr(), R, n, n_new, sdf_m.data_ptr())
            sdf = sdf_m
        z, n = z_m, n + n_new
    return z


@torch.no_grad()
def nerf_alpha(w: NerfW, pts, dirs, dists, planes):
    """alpha of compute_density_alpha (ZT:1531-1539) on a flat list of samples."""
    _, alpha, _ = nerf_forward(w, pts.contiguous(), dirs.contiguous(), dists.contiguous(), planes)
    return alpha


@torch.no_grad()
def importance_merge(z, alpha, n_new):
    """upsample_nerf + cat_z_vals_nerf (ZT:1367-1397): weights = alpha * T, sample_pdf(z, weights[:, :-1], n_new, det),
    sorted merge -- one launch of `alpha_importance_kernel` (csrc/sampling.cu) on the [R_miss, n] rows."""
    R, n = z.shape
    dev = z.device
    _, utabs = _tables(dev)
    if n_new not in utabs:
        utabs[n_new] = torch.linspace(0.5 / n_new, 1.0 - 0.5 / n_new, n_new).to(dev)
    out = _f(R, n + n_new, dev=dev)
    if R > 0:
        call("nunerf_alpha_importance", z.contiguous().float().data_ptr(), alpha.contiguous().float().data_ptr(), R, n, n_new,
             utabs[n_new].data_ptr(), out.data_ptr())
    return out


class SegCompositeFn(torch.autograd.Function):
    """Stage-2 per-segment compositing in linear colour (ZT:1942-1951) on dense [N, S] rows:
    (alpha [N,S], sRGB colour [N,S,3]) -> (rgb_lin [N,3] = sum_j w_j srgb_to_linear(c_j), t_end [N] = prod_j(1 - alpha_j + 1e-7)).
    Forward and backward are one kernel each (csrc/composite.cu); the backward recomputes the transmittance."""

    @staticmethod
    def forward(ctx, alpha, color):
        N, S = alpha.shape
        a, c = alpha.contiguous().float(), color.contiguous().float()
        rgb, t_end = _f(N, 3, dev=a.device), _f(N, dev=a.device)
        if N > 0:
            call("nunerf_seg_composite_fwd", a.data_ptr(), c.data_ptr(), N, S, rgb.data_ptr(), t_end.data_ptr())
        ctx.save_for_backward(a, c)
        return rgb, t_end

    @staticmethod
    def backward(ctx, g_rgb, g_t):
        a, c = ctx.saved_tensors
        N, S = a.shape
        da, dc = _f(N, S, dev=a.device), _f(N, S, 3, dev=a.device)
        if N > 0:
            g_rgb = (torch.zeros(N, 3, device=a.device) if g_rgb is None else g_rgb).contiguous().float()
            g_t = (torch.zeros(N, device=a.device) if g_t is None else g_t).contiguous().float()
            call("nunerf_seg_composite_bwd", a.data_ptr(), c.data_ptr(), N, S, g_rgb.data_ptr(), g_t.data_ptr(),
                 da.data_ptr(), dc.data_ptr())
        return da, dc


def _srgb_to_linear(x):
    eps = torch.finfo(torch.float32).eps
    return torch.where(x <= 0.04045, 25.0 / 323.0 * x, ((200.0 * x + 11.0) / 211.0).clamp(min=eps) ** (12.0 / 5.0))


def segment_geometry(cand):
    """Per-sample geometry of one path segment (ZT:1859-1869): cand [N, S+1, 3] sampled points, the last one being the
    surface hit (shaded separately).  Returns pts [N,S,3], Euclidean dists [N,S] (last repeated), the inner mask."""
    pts = cand[:, :-1, :].contiguous()
    d = pts[:, 1:] - pts[:, :-1]
    dists = torch.linalg.norm(d, dim=-1)
    dists = torch.cat([dists, dists[:, -1:]], -1)
    inner = torch.norm(pts, dim=-1) <= 1.0
    return pts, dists, inner


def inner_tape(pts_in, dirs_in, dists_in, cos_anneal, exp_max):
    t = CoreTape()
    t.pts_in, t.dirs_in, t.dists_in = pts_in, dirs_in, dists_in
    t.n_in, t.cos_anneal, t.exp_max = pts_in.shape[0], float(cos_anneal), float(exp_max)
    return t


def surface_forward(w1: Stage1Weights, pts, normals, dirs, exp_max):
    """AppShadingNetwork_S2.forward (field.py:909-1010) at the mesh hits: stage-1 SDF feature vector (ZT:1519-1529),
    stage-1 predictors, MESH normal, no refraction-light term.  Returns the tape (c_in = sRGB colour, trans, nov)."""
    M, dev, planes = pts.shape[0], pts.device, w1.planes
    t = CoreTape()
    t.pts_in, t.dirs_in, t.n_in, t.exp_max = pts, dirs, M, float(exp_max)
    t.xm = P(M, 320, planes, dev)
    f32_to_planes(pts, t.xm, M, 3, 64, col=256)
    t.sdf = sdf_forward(w1.sdf, pts, planes, t.xm)         # value pass fills the feature columns of xm
    shade_forward(w1, t, normals.contiguous(), no_refraction=True)
    return t


def surface_backward(w1: Stage1Weights, t, d_color, d_trans, want_geo=False, d_nov=None):
    """Reverse of surface_forward; want_geo: g["d_pts"], g["d_normals"], g["d_dirs"] (d_nov = gradient arriving at the
    NoV output, which only the geometry moves)."""
    return inner_backward(w1, t, None, d_color, None, d_trans, None, False, surface=True, want_geo=want_geo,
                          d_nov_ext=d_nov)


def shading_buffers(w: Stage1Weights, t, exp_max):
    """The intermediate results of AppShadingNetwork.forward(inter_results=True) (field.py:749-772) rebuilt from the
    predictor heads of a shade_forward tape: per-ray tensors of the eval path only (ZT:646-654, network/metrics.py)."""
    M = t.n_in
    nov, trans = t.nov[:, None], t.trans[:, None]
    tn = torch.clamp(1.0 - nov, 0.0, 1.0)
    refl_w = torch.clamp(0.04 + 0.96 * tn * tn * tn * tn * tn, 0.0, 1.0)
    head = lambda tp, n: tp.head[:, :n]
    met = torch.sigmoid(head(t.mat["metallic_predictor"], 1))
    rough = torch.sigmoid(head(t.mat["roughness_predictor"], 1))
    alb = torch.sigmoid(head(t.mat["albedo_predictor"], 3))
    ex = lambda x: torch.exp(torch.clamp(x, max=exp_max))
    lo = head(t.lo_, 3)
    diffuse_light, direct, direct0 = ex(lo[:M]), ex(lo[M:2 * M]), ex(lo[2 * M:3 * M])
    li = head(t.li_, 3)
    ind, ind0 = ex(li[:M]), ex(li[M:2 * M])
    occ = head(t.lw_, 1) * 0.5 + 0.5
    occ_c = torch.clamp(occ, 0.0, 1.0)
    light = ind * occ_c + direct * (1 - occ_c)
    light0 = ind0 * occ_c + direct0 * (1 - occ_c)
    refr = ex(head(t.lr_, 3))
    diffuse_albedo = (1 - met) * alb
    spec_albedo = 0.04 * (1 - met) + met * alb
    fg = _fg_lookup_torch(w.lut, torch.clamp(nov[:, 0], 0.0, 1.0), torch.clamp(rough[:, 0], 0.0, 1.0))
    spec_ref = spec_albedo * fg[:, 0:1] + fg[:, 1:2]
    spec_color = _lin2srgb(spec_ref * light)
    c01 = lambda x: torch.clamp(x, 0.0, 1.0)
    return {
        "specular_albedo": spec_albedo, "specular_ref": c01(spec_ref), "specular_light": c01(_lin2srgb(light0)),
        "specular_color": c01(spec_color * (1 - trans) + refl_w * light0 * trans),
        "diffuse_albedo": diffuse_albedo, "diffuse_light": c01(_lin2srgb(diffuse_light)),
        "diffuse_color": c01(_lin2srgb(diffuse_albedo * diffuse_light)),
        "metallic": met, "transmission_weight": trans, "roughness": rough, "occ_prob": c01(occ),
        "indirect_light": ind, "refraction_light": c01(_lin2srgb((1 - refl_w) * refr * trans)),
        "reflection_weight": refl_w,
    }


def surface_extras(w1: Stage1Weights, t, exp_max):
    """eval-mode buffers of field.py:981-1001, rebuilt from the predictor heads (per-ray tensors)."""
    M = t.n_in
    nov, trans = t.nov[:, None], t.trans[:, None]
    tn = torch.clamp(1.0 - nov, 0.0, 1.0)
    rw = torch.clamp(0.04 + 0.96 * tn * tn * tn * tn * tn, 0.0, 1.0)
    head = lambda tp, n: tp.head[:, :n]
    met = torch.sigmoid(head(t.mat["metallic_predictor"], 1))
    rough = torch.sigmoid(head(t.mat["roughness_predictor"], 1))
    alb = torch.sigmoid(head(t.mat["albedo_predictor"], 3))
    ex = lambda x: torch.exp(torch.clamp(x, max=exp_max))
    lo = head(t.lo_, 3)
    direct, direct0 = ex(lo[M:2 * M]), ex(lo[2 * M:3 * M])
    li = head(t.li_, 3)
    ind, ind0 = ex(li[:M]), ex(li[M:2 * M])
    occ = torch.clamp(head(t.lw_, 1) * 0.5 + 0.5, 0.0, 1.0)
    light = ind * occ + direct * (1 - occ)
    light0 = ind0 * occ + direct0 * (1 - occ)
    spec_albedo = 0.04 * (1 - met) + met * alb
    fg = _fg_lookup_torch(w1.lut, torch.clamp(nov[:, 0], 0.0, 1.0), torch.clamp(rough[:, 0], 0.0, 1.0))
    spec_ref = spec_albedo * fg[:, 0:1] + fg[:, 1:2]
    spec_color = _lin2srgb(spec_ref * light)
    return {"specular_ref": torch.clamp(spec_ref, 0.0, 1.0),
            "specular_light": torch.clamp(_lin2srgb(light0), 0.0, 1.0),
            "specular_color": torch.clamp(spec_color * (1 - trans) + rw * light0 * trans, 0.0, 1.0)}


def _lin2srgb(x):
    eps = torch.finfo(torch.float32).eps
    return torch.where(x <= 0.0031308, 323.0 / 25.0 * x, (211.0 * torch.clamp(x, min=eps) ** (5.0 / 12.0) - 11.0) / 200.0)


def _fg_lookup_torch(lut, u, v):
    """dr.texture(FG_LUT, (u, v), linear, clamp) on per-ray tensors (eval buffers only)."""
    tex = lut.reshape(256, 256, 2)
    fx = torch.clamp(u * 256 - 0.5, 0.0, 255.0)
    fy = torch.clamp(v * 256 - 0.5, 0.0, 255.0)
    x0, y0 = fx.floor().long(), fy.floor().long()
    x1, y1 = torch.clamp(x0 + 1, max=255), torch.clamp(y0 + 1, max=255)
    tx, ty = (fx - x0)[:, None], (fy - y0)[:, None]
    c00, c01, c10, c11 = tex[y0, x0], tex[y0, x1], tex[y1, x0], tex[y1, x1]
    return (c00 * (1 - tx) + c01 * tx) * (1 - ty) + (c10 * (1 - tx) + c11 * tx) * ty

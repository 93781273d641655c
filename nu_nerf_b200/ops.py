"""Thin Python wrappers over the C-ABI (nu_nerf_b200/_lib.py): plane buffers, dense layers, weight preparation.

Nothing here computes on the CPU or with torch kernels beyond allocating / zero-filling device memory.
"""
import ctypes as C
import os

import torch

from . import _lib
from ._lib import call, ptr

BF16 = torch.bfloat16
# NUNERF_GEMM_IMPL=simt selects the SIMT debug kernels (debugging aid only, never a fallback)
GEMM_IMPL = 1 if os.environ.get("NUNERF_GEMM_IMPL", "tc") == "simt" else 0
SMEM_B_BUDGET = 148224  # bytes of shared memory the resident weight tile may take (>= 3 activation stages left)


def pad(n, m):
    return (n + m - 1) // m * m


class P:
    """A bf16 plane matrix [rows, planes*w]: hi plane in columns [0,w), lo plane in [w,2w) when planes == 2."""
    __slots__ = ("t", "rows", "w", "planes", "ld", "lo", "ptr")

    def __init__(self, rows, w, planes, device, zero=False):
        assert w % 8 == 0
        self.rows, self.w, self.planes = rows, w, planes
        self.ld = planes * w
        self.lo = w if planes == 2 else 0
        self.t = (torch.zeros if zero else torch.empty)(max(rows, 1), self.ld, dtype=BF16, device=device)
        self.ptr = self.t.data_ptr()

    def at(self, row=0, col=0):
        return self.ptr + 2 * (row * self.ld + col)

    def float(self, rows=None, cols=None):
        """fp32 copy (hi + lo) -- for tests."""
        rows = self.rows if rows is None else rows
        cols = self.w if cols is None else cols
        out = torch.empty(rows, cols, dtype=torch.float32, device=self.t.device)
        call("nunerf_from_planes", self.ptr, rows, cols, self.ld, self.lo, out.data_ptr(), cols)
        return out


def to_planes(src, dst: P, rows, cols, transpose=False, scale=1.0, dst_rows=None, dst_cols=None, col_off=0, row_off=0):
    """src fp32 [rows, cols] (contiguous) -> block of dst, zero padded."""
    assert src.dtype == torch.float32 and src.is_contiguous()
    lr = cols if transpose else rows
    lc = rows if transpose else cols
    dst_rows = lr if dst_rows is None else dst_rows
    dst_cols = lc if dst_cols is None else dst_cols
    call("nunerf_to_planes", src.data_ptr(), rows, cols, src.stride(0) if src.dim() == 2 else cols, int(transpose),
         float(scale), dst.ptr, dst_rows, dst_cols, dst.ld, dst.lo, col_off, row_off)


def f32_to_planes(a, dst: P, M, C_, width, col=0, b=None):
    call("nunerf_f32_to_planes", a.data_ptr(), a.stride(0) if a.dim() == 2 else 1, ptr(b),
         (b.stride(0) if b.dim() == 2 else 1) if b is not None else 0, M, C_, width, dst.ptr, dst.ld, dst.lo, col)


def linear(A: P, B: P, M, N, K, *, a_col=0, b_row=0, bias=None, act=0, aux: P = None, aux_mode=0, aux_col=0,
           add: P = None, add_col=0, out: P = None, out_col=0, out_f32=None, n_store=0, out_scale=1.0,
           mask_in=None, mask_out=None, mask_col=0):
    """out[:, out_col:out_col+N] = epi(A[:, a_col:a_col+K] @ B[b_row:b_row+N, :K]^T); splits N so the weight tile
    stays resident in shared memory (see csrc/gemm.cu)."""
    assert K % 64 == 0 and N % 16 == 0
    planes_b = 2 if B.lo else 1
    n_tile = min(256, (SMEM_B_BUDGET // (K * 2 * planes_b)) // 16 * 16)
    assert n_tile >= 16
    if n_tile >= 64:
        n_tile = n_tile // 64 * 64      # 64-column chunks (mask words, TMA-store boxes) never straddle a split
    n_store = n_store if n_store else N
    p = _lib.LinearT()
    for n0 in range(0, N, n_tile):
        nt = min(n_tile, N - n0)
        if n0 >= n_store:
            break
        p.A, p.lda, p.a_lo_off = A.at(0, a_col), A.ld, A.lo
        p.B, p.ldb, p.b_lo_off = B.at(b_row + n0, 0), B.ld, B.lo
        p.M, p.N, p.K = M, nt, K
        p.bias = (bias.data_ptr() + 4 * n0) if bias is not None else None
        p.act = act
        if aux is not None and aux_mode in (1, 2):
            p.aux, p.ldaux, p.aux_lo_off, p.aux_mode = aux.at(0, aux_col + n0), aux.ld, aux.lo, aux_mode
        else:
            p.aux, p.ldaux, p.aux_lo_off, p.aux_mode = None, 0, 0, 0
        if mask_in is not None:
            # uint8 [M, pitch] bit masks; chunks of 64 columns = 8 bytes
            assert (mask_col + n0) % 64 == 0
            p.mask_in, p.ldmask_in, p.aux_mode = mask_in.data_ptr() + (mask_col + n0) // 8, mask_in.stride(0), 3
        else:
            p.mask_in, p.ldmask_in = None, 0
        if mask_out is not None:
            assert n0 % 64 == 0
            p.mask_out, p.ldmask_out = mask_out.data_ptr() + n0 // 8, mask_out.stride(0)
        else:
            p.mask_out, p.ldmask_out = None, 0
        if add is not None:
            p.add, p.ldadd, p.add_lo_off = add.at(0, add_col + n0), add.ld, add.lo
        else:
            p.add, p.ldadd, p.add_lo_off = None, 0, 0
        p.out_scale = out_scale
        if out is not None:
            p.out, p.ldo, p.out_lo_off = out.at(0, out_col + n0), out.ld, out.lo
        else:
            p.out, p.ldo, p.out_lo_off = None, 0, 0
        if out_f32 is not None:
            p.out_f32, p.ldo32 = out_f32.data_ptr() + 4 * n0, out_f32.stride(0)
        else:
            p.out_f32, p.ldo32 = None, 0
        p.n_store = min(nt, n_store - n0)
        p.impl = GEMM_IMPL
        call("nunerf_linear", C.byref(p))


# widest input the fused chain takes: 320 columns (5 K-blocks; the [feature | p] rows of the material predictors) with the
# default single-CTA kernel, 256 with the opt-in TS / CTA-pair variants (csrc/chain.cu)
CHAIN_K0_MAX = 256 if (os.environ.get("NUNERF_CHAIN_IMPL", "ss").startswith("t") or
                       os.environ.get("NUNERF_CHAIN_PAIR", "1") == "2" or
                       os.environ.get("NUNERF_CHAIN_K0_MAX", "") == "256") else 320


def chain(X, M, K0, layers, x_col=0, timeline=None, pts=None):
    """Fused chain of dense layers (csrc/chain.cu, bf16 single-plane operands): `layers` is a list of dicts with keys
    W (P, K-major weight), N, K and optionally n_real, bias, act, mask_out, mask_in, store (P) / store_col, out32, n32,
    keep.  See nunerf_mlp_chain_t in include/nunerf.h."""
    assert len(layers) <= 10
    a = _lib.MlpChainT()
    if X is not None:
        assert X.planes == 1
        a.x, a.ldx, a.K0 = X.at(0, x_col), X.ld, K0
    a.pts = ptr(pts)
    a.M, a.n_layers = M, len(layers)
    for l, d in enumerate(layers):
        L = a.layer[l]
        W = d["W"]
        assert W.planes == 1
        L.w, L.ldw, L.N, L.K = W.at(d.get("w_row", 0), 0), W.ld, d["N"], d["K"]
        L.n_real = d.get("n_real", 0)
        L.bias = ptr(d.get("bias"))
        L.act = d.get("act", 0)
        mo, mi = d.get("mask_out"), d.get("mask_in")
        L.mask_out, L.ldmask_out = (mo.data_ptr(), mo.stride(0)) if mo is not None else (None, 0)
        L.mask_in, L.ldmask_in = (mi.data_ptr(), mi.stride(0)) if mi is not None else (None, 0)
        st = d.get("store")
        L.store, L.ld_store = (st.at(0, d.get("store_col", 0)), st.ld) if st is not None else (None, 0)
        o32 = d.get("out32")
        L.out32, L.ldo32, L.n32 = (o32.data_ptr(), o32.stride(0), d.get("n32", d["N"])) if o32 is not None else (None, 0, 0)
        L.keep = int(d.get("keep", 0))
        L.mask_perm = int(d.get("mask_perm", 0))
        L.cat_pe, L.aux_mode = int(d.get("cat_pe", 0)), int(d.get("aux_mode", 0))
        for k_, f_, l_ in (("aux1", "aux1", "ld_aux1"), ("aux2", "aux2", "ld_aux2"), ("e_out", "e_out", "ld_e")):
            pl = d.get(k_)
            if pl is not None:
                assert pl.planes == 1
                setattr(L, f_, pl.ptr)
                setattr(L, l_, pl.ld)
    a.timeline = ptr(timeline)
    call("nunerf_mlp_chain", C.byref(a))


def linear_dw(dZ: P, X: P, M, N, K, dW, *, z_col=0, x_col=0, db=None):
    """dW[:N, :K] += dZ[:, z_col:z_col+N]^T @ X[:, x_col:x_col+K]   (dW fp32, pre-zeroed / accumulating);
    db[:N] += column sums of dZ (the bias gradient) in the same launch when given."""
    assert K % 64 == 0
    p = _lib.DwT()
    p.dZ, p.ldz, p.z_lo_off = dZ.at(0, z_col), dZ.ld, dZ.lo
    p.X, p.ldx, p.x_lo_off = X.at(0, x_col), X.ld, X.lo
    p.M, p.N, p.K = M, N, K
    p.dW, p.lddw = dW.data_ptr(), dW.stride(0)
    p.impl = GEMM_IMPL
    p.db = None
    if db is not None:
        if GEMM_IMPL == 1:
            colsum(dZ, M, N, db, z_col=z_col)
        else:
            p.db = db.data_ptr()
    call("nunerf_linear_dw", C.byref(p))


def colsum(Z: P, M, N, out, z_col=0):
    call("nunerf_colsum", Z.at(0, z_col), Z.ld, Z.lo, M, N, out.data_ptr())

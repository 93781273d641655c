"""Persistent tensor-core operands of every dense layer + the two-launch weight pipeline (csrc/weights.cu).

A `Dense` is one *prepared* matrix: a (possibly rotated / row-sliced / scaled) view of a source parameter
(weight-normed `weight_v, weight_g` or a plain `weight`) materialised every step as bf16 planes
  Wk  [pad16(rows), planes * pad64(K)]   K-major   (forward  Y = X W^T)
  WTk [pad64(K), planes * pad64(rows)]   K-major   (backward dX = dZ W)
plus a zero-padded bias copy, and its fp32 gradient buffers dW [pad16(rows), pad64(K)], db [pad16(rows)].
`WeightBank.prepare()` refreshes all of them in ONE kernel launch; `WeightBank.backward()` turns the accumulated
dW / db into gradients of the source parameters in ONE launch, added in place to `param.grad`.
"""
import torch

from . import _lib
from ._lib import call
from .ops import P, pad


class Dense:
    def __init__(self, bank, v, g=None, bias=None, *, rows=None, row_rot=0, col_rot=0, scale=1.0, need_k=True,
                 need_t=True, grad=True, row_f32=False, t_cols=None, wtk_share=None):
        self.v, self.g, self.bias = v, g, bias
        self.srcN, self.K = v.shape
        r0, r1 = rows if rows is not None else (0, self.srcN)
        self.row0, self.N = r0, r1 - r0
        self.row_rot, self.col_rot, self.scale = row_rot, col_rot, float(scale)
        self.Np, self.Kp, self.Np64 = pad(self.N, 16), pad(self.K, 64), pad(self.N, 64)
        dev, planes = v.device, bank.planes
        self.Wk = P(self.Np, self.Kp, planes, dev, zero=True) if need_k else None
        self.WTk = P(self.Kp, t_cols if t_cols else self.Np64, planes, dev, zero=True) if need_t else None
        self.wtk_col_off = 0
        if wtk_share is not None:           # write the transposed rows into another Dense's WTk at a column offset
            self.WTk, self.wtk_col_off = wtk_share
        self.b = torch.zeros(self.Np, device=dev) if bias is not None else None
        self.has_grad = grad
        self.dW = self.db = None            # views into the bank's flat gradient buffer (finalize)
        self.row_f32 = torch.zeros(self.N, self.K, device=dev) if row_f32 else None
        bank.denses.append(self)


class WeightBank:
    def __init__(self, planes, device):
        self.planes, self.device = planes, device
        self.denses = []
        self._final = False
        self._grad_ptrs = None

    # ---------------------------------------------------------------------------------------------- build
    def finalize(self):
        dev = self.device
        n_dw = sum(d.Np * d.Kp for d in self.denses if d.has_grad)
        n_db = sum(d.Np for d in self.denses if d.has_grad)
        self.gflat = torch.zeros(n_dw + n_db, device=dev)
        off = 0
        for d in self.denses:
            if d.has_grad:
                d.dW = self.gflat[off:off + d.Np * d.Kp].view(d.Np, d.Kp)
                off += d.Np * d.Kp
        for d in self.denses:
            if d.has_grad:
                d.db = self.gflat[off:off + d.Np]
                off += d.Np
        # 1/|v| per weight-normed source parameter (shared by every Dense derived from it)
        self._inv_norm = {}
        for d in self.denses:
            if d.g is not None and d.v.data_ptr() not in self._inv_norm:
                self._inv_norm[d.v.data_ptr()] = torch.zeros(d.srcN, device=dev)
        blk_desc, blk_row = [], []
        for i, d in enumerate(self.denses):
            blk_desc += [i] * d.N
            blk_row += list(range(d.N))
        self.n_blocks = len(blk_desc)
        self.blk_desc = torch.tensor(blk_desc, dtype=torch.int32, device=dev)
        self.blk_row = torch.tensor(blk_row, dtype=torch.int32, device=dev)
        self._host = (_lib.WDesc * len(self.denses))()
        self._fill_static()
        self._descs = None
        self._upload()
        self._final = True

    def _fill_static(self):
        for i, d in enumerate(self.denses):
            h = self._host[i]
            h.v, h.g = d.v.data_ptr(), (d.g.data_ptr() if d.g is not None else None)
            h.N, h.K, h.ld, h.scale = d.srcN, d.K, d.v.stride(0), d.scale
            h.row_rot, h.col_rot, h.src_row0, h.n_rows = d.row_rot, d.col_rot, d.row0, d.N
            if d.Wk is not None:
                h.wk, h.wk_ld, h.wk_lo, h.wk_row_off = d.Wk.ptr, d.Wk.ld, d.Wk.lo, 0
            if d.WTk is not None:
                h.wtk, h.wtk_ld, h.wtk_lo, h.wtk_col_off = d.WTk.ptr, d.WTk.ld, d.WTk.lo, d.wtk_col_off
            h.inv_norm = self._inv_norm[d.v.data_ptr()].data_ptr() if d.g is not None else None
            h.row_f32 = d.row_f32.data_ptr() if d.row_f32 is not None else None
            if d.bias is not None:
                h.bias_src, h.bias_dst = d.bias.data_ptr(), d.b.data_ptr()
            if d.has_grad:
                h.dW, h.lddw, h.dw_row_off, h.db = d.dW.data_ptr(), d.Kp, 0, d.db.data_ptr()

    def _upload(self):
        """Descriptor table -> device through one of two pinned staging buffers (asynchronous copy: a pageable copy would
        stall the host until the stream drains, in the middle of a step)."""
        raw = bytes(self._host)
        if torch.device(self.device).type != "cuda":       # (host-side construction in CPU-only tests)
            self._descs = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(self.device)
            return
        if self._descs is None:
            self._descs = torch.empty(len(raw), dtype=torch.uint8, device=self.device)
            self._pin = [torch.empty(len(raw), dtype=torch.uint8).pin_memory() for _ in range(2)]
            self._pin_ev = [None, None]
            self._pin_i = 0
        i = self._pin_i
        self._pin_i ^= 1
        if self._pin_ev[i] is not None:
            self._pin_ev[i].synchronize()                  # the copy that last used this staging buffer has completed
        self._pin[i].copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
        self._descs.copy_(self._pin[i], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._pin_ev[i] = ev

    def source_ptrs(self):
        return tuple(d.v.data_ptr() for d in self.denses)

    # ---------------------------------------------------------------------------------------------- per step
    def prepare(self):
        """One launch: weight-norm + layout + bf16 planes + bias copies for every layer."""
        call("nunerf_weights_prepare", self._descs.data_ptr(), self.blk_desc.data_ptr(), self.blk_row.data_ptr(),
             self.n_blocks)

    def zero_grads(self):
        self.gflat.zero_()

    def backward(self):
        """One launch: accumulated dW / db -> += into the source parameters' .grad (allocated if missing).
        Parameters with requires_grad=False are skipped (null gradient pointers in the descriptor table).

        Autograd contract of the engine's nodes (renderer_zerothick._RenderCoreFn / _SdfValueFn, renderer_stage2._NerfFn /
        _InnerFn / _SurfaceFn): parameter gradients are ADDED to .grad here, from inside backward, and the nodes return
        None for their parameter inputs.  Supported: loss.backward() with any optimiser reading .grad (what the
        reference trainer does).  Not supported: torch.autograd.grad() w.r.t. parameters, per-parameter grad hooks /
        DistributedDataParallel bucket hooks (use nu_nerf_b200.dist for data parallelism), retain_graph double
        backward.  The .grad tensors allocated here are views of one pooled buffer reused across steps: a caller that
        keeps last step's .grad after zero_grad(set_to_none=True) must clone it."""
        ptrs = []
        # parameters without a .grad (first step, or after zero_grad(set_to_none=True)) get views of ONE zero-filled
        # buffer: one fill launch instead of one per tensor (~300 for a stage-2 renderer)
        missing, seen = [], set()
        for d in self.denses:
            if d.has_grad:
                for p in (d.v, d.g, d.bias):
                    if p is not None and p.requires_grad and p.grad is None and id(p) not in seen:
                        seen.add(id(p))
                        missing.append(p)
        if missing:
            sizes = [(p.numel() + 3) // 4 * 4 for p in missing]          # keep every view 16-byte aligned
            key = tuple(id(p) for p in missing)
            pool = getattr(self, "_grad_pool", None)
            if pool is not None and pool[0] == key:
                # (a caller that keeps last step's .grad tensors after zero_grad(set_to_none=True) must clone them)
                flat = pool[1]                                           # same parameters as last time: same storage, so
                flat.zero_()                                             # the descriptor table needs no new upload
            else:
                flat = torch.zeros(sum(sizes), dtype=missing[0].dtype, device=missing[0].device)
                self._grad_pool = (key, flat)
            off = 0
            for p, n in zip(missing, sizes):
                p.grad = flat[off:off + p.numel()].view_as(p)
                off += n
        for d in self.denses:
            if not d.has_grad:
                continue
            gp = lambda p: p.grad.data_ptr() if (p is not None and p.requires_grad) else 0
            ptrs.append((gp(d.v), gp(d.g), gp(d.bias)))
        ptrs = tuple(ptrs)
        if ptrs != self._grad_ptrs:
            j = 0
            for i, d in enumerate(self.denses):
                if not d.has_grad:
                    continue
                h = self._host[i]
                dv, dg, dbias = ptrs[j]
                j += 1
                h.dv, h.dg, h.dbias = (dv or None), (dg or None), (dbias or None)
                if d.bias is None:
                    h.db = None
            self._upload()
            self._grad_ptrs = ptrs
        call("nunerf_weights_backward", self._descs.data_ptr(), self.blk_desc.data_ptr(), self.blk_row.data_ptr(),
             self.n_blocks)

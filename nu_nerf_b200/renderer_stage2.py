"""Stage2Renderer -- the zero-thickness nested-refraction renderer of network/renderer_zerothick.py:868-2011 ("ZT")
on the sm_100a engine: forward (ray_trace + render_core, train- and eval-mode outputs) and the backward of the trainer
loss with respect to every parameter: the FIELD parameters (stage-1 NeRF++ / SDF feature / predictors at the surface
hits, inner SDF, inner shading, inner variance) and IORs_pred, whose gradient flows through the refracted sample
positions into every field's input.

  ray_trace   ZT:1571-1828   <= 3 bounces: BVH closest hit + re-intersection (csrc/bvh.cu), IoR MLP on tensor cores,
                             Snell / TIR kernel, segment sampling (256 uniform | 64 + 2 x 32 SDF-guided with the warp
                             per-ray up-sampling kernel | 192 + 64 NeRF-guided on [0.1, 64])
  render_core ZT:1835-2011   per segment: NeRF++ on the outer samples, inner SDF + shading on segment 1, surface
                             shading at the hit with the mesh normal and the stage-1 predictors, linear-space
                             compositing with the throughput chain T *= T_end (1 - schlick) transmission
  _replay_geometry           the trace again as a differentiable function of IORs_pred: nodes with the trace's own
                             values in the forward and reverse kernels in the backward (hit_interp_bwd, refract_bounce_bwd,
                             points_bwd); the IoR MLP (<= R rows per bounce) is the only torch-autograd part

Each field evaluation is one autograd node (forward = explicit launch sequence, backward = hand-derived reverse sequence
of engine.py, gradients added in place to .grad by the WeightBank).  When the geometry carries a graph the same nodes
also return the gradient with respect to their sample positions / directions / interval lengths / normals
(position-gradient kernels: PE backward, PE Hessian of the SDF gradient, NeRF++ inverted-sphere backward, csrc/field.cu).
The per-ray / per-segment bookkeeping between the nodes (mask compaction, throughput chain, reverse scatter-add) is torch
indexing, as in the reference.  Lists returned by ray_trace have the reference's exact structure (per-segment compacted
rows in boolean-mask order), so render_core accepts the reference's ray_trace output and vice versa.
cfg['frozen_ior'] = True keeps the geometry constant (no replay, no position gradients: IORs_pred gets no gradient).
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .field import (SDFNetwork, SingleVarianceNetwork, NeRFNetwork, AppShadingNetwork, AppShadingNetwork_S2, IoRNetwork,
                    ThicknessNetwork)


def _engine():
    from . import engine
    return engine


def srgb_to_linear(x):
    """utils/raw_utils.py:21-27."""
    eps = torch.finfo(torch.float32).eps
    return torch.where(x <= 0.04045, 25.0 / 323.0 * x, ((200.0 * x + 11.0) / 211.0).clamp(min=eps) ** (12.0 / 5.0))


def linear_to_srgb(x):
    eps = torch.finfo(torch.float32).eps
    return torch.where(x <= 0.0031308, 323.0 / 25.0 * x, (211.0 * torch.clamp(x, min=eps) ** (5.0 / 12.0) - 11.0) / 200.0)


class _StraightThrough(torch.autograd.Function):
    """value of `traced` (the no-grad kernel trace), gradient to `replayed` (the same quantity as a differentiable
    expression): no arithmetic in the forward, the incoming gradient is handed through in the backward."""

    @staticmethod
    def forward(ctx, traced, replayed):
        return traced.view_as(traced)

    @staticmethod
    def backward(ctx, g):
        return None, g


class _HitFn(torch.autograd.Function):
    """Scene.Dintersect on the recorded triangles as a differentiable node (DiffRender.py:61-124): (o, d) of the rays
    that hit -> hit point x = o + t d and the signed, interpolated unit normal.  Forward values are the trace's own when
    given (else `nunerf_hit_interp` is run again); backward = `hit_interp_bwd_kernel`."""

    @staticmethod
    def forward(ctx, o_c, d_c, pack):
        eng = _engine()
        scene, tri, inside, x, n = pack
        o_c, d_c = o_c.contiguous(), d_c.contiguous()
        N = o_c.shape[0]
        if x is None or n is None:
            uvt, x, n = (torch.empty(N, 3, device=o_c.device) for _ in range(3))
            if N > 0:
                eng.call("nunerf_hit_interp", scene.optix_mesh.bvh.tri_verts.data_ptr(), scene.tri_normals.data_ptr(),
                         tri.data_ptr(), o_c.data_ptr(), d_c.data_ptr(), N, uvt.data_ptr(), x.data_ptr(), n.data_ptr())
            n = -n if inside else n
        ctx.save_for_backward(o_c, d_c)
        ctx.pack = (scene, tri, inside)
        return x.view_as(x), n.view_as(n)

    @staticmethod
    def backward(ctx, g_x, g_n):
        eng = _engine()
        o_c, d_c = ctx.saved_tensors
        scene, tri, inside = ctx.pack
        N = o_c.shape[0]
        g_o, g_d = torch.empty_like(o_c), torch.empty_like(d_c)
        zero = lambda g: torch.zeros(N, 3, device=o_c.device) if g is None else g.contiguous().float()
        if N > 0:
            eng.call("nunerf_hit_interp_bwd", scene.optix_mesh.bvh.tri_verts.data_ptr(), scene.tri_normals.data_ptr(),
                     tri.data_ptr(), o_c.data_ptr(), d_c.data_ptr(), N, int(inside), zero(g_x).data_ptr(),
                     zero(g_n).data_ptr(), g_o.data_ptr(), g_d.data_ptr())
        return g_o, g_d, None


class _RefractFn(torch.autograd.Function):
    """Snell step of the zero-thickness bounce (ZT:1655-1684) on the rays that passed the TIR test: (x, signed normal,
    incoming direction, effective ratio) -> (next origin, next direction).  Forward values are the trace's own when
    given; backward = `refract_bounce_bwd_kernel`."""

    @staticmethod
    def forward(ctx, x, n, d, eta, pack):
        o_next, d_next = pack
        n, d, eta = n.contiguous(), d.contiguous(), eta.contiguous()
        if o_next is None or d_next is None:
            cos_i = -(n * d).sum(-1, keepdim=True)
            e = eta.reshape(-1, 1)
            d_tmp = e * d + (e * cos_i - torch.sqrt(1.0 - (1.0 - cos_i * cos_i) * e * e)) * n
            o_next = x + d_tmp * 1e-5
            d_next = d_tmp / (torch.linalg.norm(d_tmp, dim=-1, keepdim=True) + 0.0001)
        ctx.save_for_backward(n, d, eta)
        return o_next.view_as(o_next), d_next.view_as(d_next)

    @staticmethod
    def backward(ctx, g_o, g_d):
        eng = _engine()
        n, d, eta = ctx.saved_tensors
        N = n.shape[0]
        zero = lambda g: torch.zeros(N, 3, device=n.device) if g is None else g.contiguous().float()
        g_x, g_n, g_dd, g_eta = torch.empty_like(n), torch.empty_like(n), torch.empty_like(n), torch.empty(N, device=n.device)
        if N > 0:
            eng.call("nunerf_refract_bounce_bwd", n.data_ptr(), d.data_ptr(), eta.data_ptr(), N, zero(g_o).data_ptr(),
                     zero(g_d).data_ptr(), g_x.data_ptr(), g_n.data_ptr(), g_dd.data_ptr(), g_eta.data_ptr())
        return g_x, g_n, g_dd, g_eta.view_as(eta), None


class _SegPointsFn(torch.autograd.Function):
    """Path points start + delta * z of one segment (ZT:1727-1731, :1760, :1799) with the values of the trace and the
    reverse reduction over the samples as a kernel (`points_bwd_kernel`)."""

    @staticmethod
    def forward(ctx, start, delta, z, traced):
        ctx.save_for_backward(z)
        return traced.view_as(traced)

    @staticmethod
    def backward(ctx, g):
        eng = _engine()
        (z,) = ctx.saved_tensors
        R, n = z.shape
        g_s, g_d = torch.empty(R, 3, device=z.device), torch.empty(R, 3, device=z.device)
        if R > 0:
            eng.call("nunerf_points_bwd", g.contiguous().float().data_ptr(), z.contiguous().data_ptr(), R, n,
                     g_s.data_ptr(), g_d.data_ptr())
        return g_s, g_d, None, None


def starts_k(rec, k, what):
    """start / direction of segment k as the (no-grad) trace produced them."""
    return rec["segments"][k][what]


def _bank_params(w):
    ps = [p for d in w.bank.denses if d.has_grad for p in (d.v, d.g, d.bias) if p is not None]
    return list({id(p): p for p in ps}.values())


def _zf(g, dev, *shape):
    return torch.zeros(*shape, device=dev) if g is None else g.contiguous().float()


class _NerfFn(torch.autograd.Function):
    """compute_density_alpha (ZT:1531-1539) on a flat list of outer samples with the stage-1 NeRF++.  pts / dirs / dists
    are differentiable inputs: when the path geometry carries a graph (IORs_pred is being trained) the backward also runs
    the position-gradient kernels (engine.nerf_backward(want_geo=True))."""

    @staticmethod
    def forward(ctx, w, pts, dirs, dists, *params):
        eng = _engine()
        tape, alpha, color = eng.nerf_forward(w.nerf, pts, dirs, dists, w.planes)
        ctx.w, ctx.tape, ctx.n = w, tape, len(params)
        return alpha, color

    @staticmethod
    def backward(ctx, d_alpha, d_color):
        eng = _engine()
        w, t = ctx.w, ctx.tape
        dev = t.dists.device
        geo = any(ctx.needs_input_grad[1:4])
        w.bank.zero_grads()
        g = eng.nerf_backward(w.nerf, t, _zf(d_alpha, dev, t.M), _zf(d_color, dev, t.M, 3), w.planes, want_geo=geo)
        w.bank.backward()
        ctx.tape = None
        d_pts, d_dirs, d_dists = g if geo else (None, None, None)
        return (None, d_pts, d_dirs, d_dists) + (None,) * ctx.n


class _InnerFn(torch.autograd.Function):
    """compute_sdf_alpha + color_network_inner (ZT:1887-1906) on the compact inner samples of segment 1."""

    @staticmethod
    def forward(ctx, pack, inv_s, pts, dirs, dists, *params):
        eng = _engine()
        w, cos_anneal, exp_max, want_inv_s = pack
        t = eng.inner_tape(pts, dirs, dists, cos_anneal, exp_max)
        eng.inner_forward(w, t)
        ctx.w, ctx.tape, ctx.n, ctx.want_inv_s = w, t, len(params), want_inv_s
        return t.a_in, t.c_in, t.gerr

    @staticmethod
    def backward(ctx, d_alpha, d_color, d_gerr):
        eng = _engine()
        w, t = ctx.w, ctx.tape
        dev, M = t.pts_in.device, t.n_in
        geo = any(ctx.needs_input_grad[2:5])
        w.bank.zero_grads()
        g = eng.inner_backward(w, t, _zf(d_alpha, dev, M), _zf(d_color, dev, M, 3), _zf(d_gerr, dev, M), None, None,
                               ctx.want_inv_s, want_geo=geo)
        w.bank.backward()
        ctx.tape = None
        d_inv = g["inv_s"].reshape(()) if ctx.want_inv_s and "inv_s" in g else None
        return (None, d_inv, g.get("d_pts"), g.get("d_dirs"), g.get("d_dists")) + (None,) * ctx.n


class _SurfaceFn(torch.autograd.Function):
    """AppShadingNetwork_S2 at the mesh hits (ZT:1908-1925): sRGB colour, transmission weight, NoV.  The hit points,
    mesh normals and incoming directions are differentiable inputs (they move with IORs_pred from the second hit on)."""

    @staticmethod
    def forward(ctx, pack, pts, normals, dirs, *params):
        eng = _engine()
        w, exp_max, holder = pack
        t = eng.surface_forward(w, pts, normals, dirs, exp_max)
        holder["tape"] = t                      # eval-mode extras are rebuilt from the predictor heads
        ctx.w, ctx.tape, ctx.n = w, t, len(params)
        ctx.geo = any(ctx.needs_input_grad[1:4])
        if not ctx.geo:
            ctx.mark_non_differentiable(t.nov)  # NoV only depends on the geometry
        return t.c_in, t.trans, t.nov

    @staticmethod
    def backward(ctx, d_color, d_trans, d_nov):
        eng = _engine()
        w, t = ctx.w, ctx.tape
        dev, M = t.pts_in.device, t.n_in
        w.bank.zero_grads()
        g = eng.surface_backward(w, t, _zf(d_color, dev, M, 3), _zf(d_trans, dev, M), want_geo=ctx.geo,
                                 d_nov=_zf(d_nov, dev, M) if ctx.geo else None)
        w.bank.backward()
        ctx.tape = None
        return (None, g.get("d_pts"), g.get("d_normals"), g.get("d_dirs")) + (None,) * ctx.n


class _InnerField:
    """The (sdf_network_inner, deviation_network_inner, color_network_inner) triple seen as a stage-1 field."""

    def __init__(self, r):
        self.sdf_network, self.deviation_network = r.sdf_network_inner, r.deviation_network_inner
        # the NeRF++ slot of the field is never evaluated for the inner field; the non-zero-thickness renderer has no
        # stage-2 level outer_nerf module (network/renderer.py:975-1017), there the slot holds the stage-1 one
        self.color_network = r.color_network_inner
        self.outer_nerf = r.outer_nerf if hasattr(r, "outer_nerf") else r.stage1_network.outer_nerf


class Stage2Renderer(nn.Module):
    default_cfg = {
        "std_net": "default", "std_act": "exp", "inv_s_init": 0.3, "freeze_inv_s_step": None,
        "sdf_net": "default", "sdf_activation": "none", "sdf_bias": 0.5, "sdf_n_layers": 8, "sdf_freq": 6,
        "sdf_d_out": 257, "geometry_init": True, "shader_config": {},
        "n_samples": 64, "n_bg_samples": 32, "inf_far": 1000.0, "n_importance": 64, "up_sample_steps": 4,
        "perturb": 1.0, "anneal_end": 50000, "train_ray_num": 1024, "test_ray_num": 1024,
        "clip_sample_variance": True, "database_name": "nerf_synthetic/lego/black_800", "is_nerf": False,
        "test_downsample_ratio": True, "downsample_ratio": 0.5, "val_geometry": False,
        "rgb_loss": "charbonier", "apply_occ_loss": True, "occ_loss_step": 20000, "occ_loss_max_pn": 2048,
        "occ_sdf_thresh": 0.01, "fixed_camera": False,
        "precision": "split",
    }

    def __init__(self, cfg, training=True):
        super().__init__()
        from .renderer_zerothick import NeROShapeRenderer, load_cfg
        from .tracer import Scene
        self.cfg = {**self.default_cfg, **cfg}
        self.is_nerf = self.cfg["is_nerf"]
        if (self.cfg["sdf_n_layers"], self.cfg["sdf_freq"], self.cfg["sdf_d_out"]) != (8, 6, 257) or \
                not self.cfg["clip_sample_variance"]:
            raise NotImplementedError("the B200 engine is built for the 8x256 / PE-6 / 257-output SDF network")
        # construction order = the reference's (ZT:919-975): same seed => bit-identical initial parameters
        self.nerf_network = NeRFNetwork(D=8, d_in=4, d_in_view=3, W=256, multires=10, multires_view=4, skips=(4,))
        self.IORs = nn.Parameter(torch.zeros(10))
        cfg1 = cfg["stage1_cfg_dir"]
        cfg1 = dict(cfg1) if isinstance(cfg1, dict) else load_cfg(cfg1)
        cfg1.setdefault("precision", self.cfg["precision"])
        self.stage1_network = NeROShapeRenderer(cfg1, training=False)
        ckpt = cfg["stage1_ckpt_dir"]
        ckpt = ckpt if isinstance(ckpt, dict) else torch.load(ckpt, map_location="cpu")
        self.stage1_network.load_state_dict(ckpt["network_state_dict"], strict=False)
        self._mesh = cfg["stage1_mesh_dir"]          # path (.ply / .npz) or (vertices, faces)
        self.scene = None                            # built on first use, on the parameters' device
        self._scene_cls = Scene
        self.IORs_pred = IoRNetwork()
        self.IoRint_pred = IoRNetwork()
        self.thickness_pred = ThicknessNetwork()
        self.outer_nerf = NeRFNetwork(D=8, d_in=4, d_in_view=3, W=256, multires=10, multires_view=4, skips=(4,))
        self.color_network = AppShadingNetwork_S2(self.cfg["shader_config"], self.stage1_network)
        self.sdf_network_inner = SDFNetwork(d_out=257, d_in=3, d_hidden=256, n_layers=8, skip_in=(4,), multires=6,
                                            bias=self.cfg["sdf_bias"], scale=1.0,
                                            geometric_init=self.cfg["geometry_init"])
        self.deviation_network_inner = SingleVarianceNetwork(init_val=self.cfg["inv_s_init"],
                                                             activation=self.cfg["std_act"])
        self.color_network_inner = AppShadingNetwork(self.cfg["shader_config"])
        self.sdf_network_inner._query = self._sdf_inner_query
        self.ray_source = None

    # ------------------------------------------------------------------ helpers identical to the reference
    def set_ray_source(self, fn):
        self.ray_source = fn

    def get_anneal_val(self, step):
        if self.cfg["anneal_end"] < 0:
            return 1.0
        return np.min([1.0, step / self.cfg["anneal_end"]])

    def compute_rgb_loss(self, rgb_pr, rgb_gt):
        if self.cfg["rgb_loss"] == "charbonier":
            return torch.sqrt(torch.sum((rgb_gt - rgb_pr) ** 2, dim=-1) + 0.001)
        raise NotImplementedError

    # ------------------------------------------------------------------ engine operands
    def _planes(self):
        return 2 if self.cfg["precision"] == "split" else 1

    def _prepare(self):
        dev = self.deviation_network_inner.variance.device
        if not torch.cuda.is_available() or dev.type != "cuda":
            raise RuntimeError("nu_nerf_b200 renders on a CUDA device only (no CPU fallback): move the module with .cuda()")
        from ._lib import require_current_device
        require_current_device(self.deviation_network_inner.variance)       # the C entry points launch on the current device's stream
        eng = _engine()
        self.stage1_network.cfg["precision"] = self.cfg["precision"]
        w1 = self.stage1_network._prepare()
        key = (self._planes(),) + tuple(p.data_ptr() for p in self.parameters())
        if getattr(self, "_w", None) is None or self._w_key != key:
            inner = eng.Stage1Weights(_InnerField(self), self._planes(), dev)
            ior = eng.IorWeights(self.IORs_pred, self._planes(), dev)
            self._w, self._w_key = (inner, ior), key
        self._w[0].refresh()
        self._w[1].refresh()
        if self.scene is None:
            m = self._mesh
            self.scene = self._scene_cls(m, device=dev) if isinstance(m, str) else self._scene_cls(m[0], m[1], device=dev)
        return w1, self._w[0], self._w[1]

    @torch.no_grad()
    def _sdf_inner_query(self, x):
        eng = _engine()
        _, wi, _ = self._prepare()
        shape = x.shape[:-1]
        return eng.sdf_infer(wi.sdf, x.reshape(-1, 3).float().contiguous(), wi.planes).reshape(*shape, 1)

    # ------------------------------------------------------------------ ZT:1571-1828
    @torch.no_grad()
    def ray_trace(self, rays_o, rays_d, prepared=None, trace=None, rec=None):
        """rec (dict): filled with the discrete decisions and per-sample parameters of the trace (hit / pass index lists,
        triangle ids, sample parameters z) for _replay_geometry, which rebuilds the same geometry as a differentiable
        function of IORs_pred."""
        eng = _engine()
        w1, wi, wior = prepared if prepared is not None else self._prepare()
        dev = rays_o.device
        o, d = rays_o.float().contiguous(), rays_d.float().contiguous()
        next_start, next_dir = o, d
        starts, directions = [o], [d]
        intersections, converges, infinity_bkgr, ior_ratios, gradient_mesh, tirs = [], [], [], [], [], []
        hit_idxs, conv_idxs = [], []
        inside = False
        for i in range(3):
            N = next_start.shape[0]
            info, hit = self.scene.Dintersect(next_start, next_dir)       # closest hit + re-intersection (a14, a15)
            if trace is not None:
                trace[f"trace_hit_{i}"], trace[f"trace_tri_{i}"] = hit.float(), info["faces_ind"]
            converged = hit.reshape(-1, 1)
            # index tensors instead of boolean-mask indexing: ONE host sync (nonzero) per mask instead of one per use
            hit_idx = hit.nonzero().squeeze(1)
            x_c = info["x"][hit_idx].contiguous()
            n_raw = info["n"][hit_idx].contiguous()
            n_c = F.normalize(n_raw, dim=-1)
            if inside:
                n_c = -n_c
            d_c = next_dir[hit_idx].contiguous()
            infinity_bkgr.append(~converged)
            M = x_c.shape[0]
            if M > 0:
                eta = 1.0 / (eng.ior_forward(wior, x_c) + 1.0)                 # ZT:1642-1643
                d_out, o_out = torch.empty(M, 3, device=dev), torch.empty(M, 3, device=dev)
                ok = torch.empty(M, dtype=torch.uint8, device=dev)
                tri_c = info["faces_ind"][hit_idx].contiguous()
                eng.call("nunerf_refract_bounce", x_c.data_ptr(), n_raw.data_ptr(),
                         d_c.data_ptr(), eta.contiguous().data_ptr(), tri_c.data_ptr(), M, int(inside),
                         d_out.data_ptr(), o_out.data_ptr(), ok.data_ptr())
                ok = ok.bool()
                ratio = (1.0 / eta if inside else eta).reshape(-1, 1)
            else:
                d_out = o_out = torch.zeros(0, 3, device=dev)
                ok = torch.zeros(0, dtype=torch.bool, device=dev)
                ratio = torch.zeros(0, 1, device=dev)
            ok_idx = ok.nonzero().squeeze(1)
            converged_out = converged.clone()
            converged_out[hit_idx] = ok.reshape(-1, 1)
            tir = torch.ones(N, 1, dtype=torch.bool, device=dev)
            tir[hit_idx] = ok.reshape(-1, 1)
            tirs.append(tir)
            next_dir, next_start = d_out[ok_idx].contiguous(), o_out[ok_idx].contiguous()
            directions.append(next_dir)
            starts.append(next_start)
            converges.append(converged_out)
            intersections.append(x_c)
            hit_idxs.append(hit_idx)
            conv_idxs.append(hit_idx[ok_idx])                 # rays of this segment that continue into the next one
            if rec is not None:
                rec.setdefault("bounces", []).append(dict(hit_idx=hit_idx, ok_idx=ok_idx, inside=inside, x=x_c, n=n_c,
                                                          tri=info["faces_ind"][hit_idx].long(),
                                                          eta=eta if M > 0 else torch.zeros(0, device=dev)))
            if ok_idx.numel() == 0:
                break
            gradient_mesh.append(n_c[ok_idx])
            ior_ratios.append(ratio[ok_idx])
            inside = not inside
        for i in range(len(tirs) - 1, 0, -1):
            m = conv_idxs[i - 1]
            tirs[i - 1][m] = tirs[i - 1][m] & tirs[i]
        # ---- per-segment sample generation (ZT:1719-1813)
        pathes = []
        for k in range(len(converges)):
            start, dk = starts[k], directions[k]
            h_idx = hit_idxs[k]                                   # rays of the segment that hit the mesh (~bk)
            n_hit, n_seg = h_idx.numel(), start.shape[0]
            end = start + dk * 4.5
            if n_hit > 0:
                end[h_idx] = intersections[k]
            n_pts = 256 if k != 1 else 128
            lin = torch.linspace(0, 1, n_pts, device=dev)
            pts = eng.segment_points(start, end - start, lin.unsqueeze(0).expand(start.shape[0], n_pts).contiguous())
            Z = lin.unsqueeze(0).repeat(n_seg, 1) if rec is not None else None
            m_idx = None
            if k == 1 and n_hit > 0:
                # inside the outer mesh: 64 uniform samples to the hit, 2 rounds of SDF-guided up-sampling with 32 new
                # samples each on the inner field; z is the unit parameter of the segment while the SDF is queried at
                # start + dir * z -- the reference's mixed parametrisation, reproduced as written (ZT:1742-1760)
                s_h, e_h, d_h = start[h_idx].contiguous(), end[h_idx].contiguous(), dk[h_idx].contiguous()
                Rh = s_h.shape[0]
                z = torch.linspace(0, 1, 64, device=dev).unsqueeze(0).expand(Rh, 64).contiguous()
                p64 = eng.segment_points(s_h, e_h - s_h, z)
                sdf = eng.sdf_infer(wi.sdf, p64.reshape(-1, 3), wi.planes).reshape(Rh, 64).contiguous()
                z = eng.upsample_rounds(wi, s_h, d_h, z, sdf, n_new=32, rounds=2)
                pts[h_idx] = eng.segment_points(s_h, e_h - s_h, z)
                if Z is not None:
                    Z[h_idx] = z
            if k != 1 and n_hit < n_seg:
                # rays that leave the scene: 192 samples on [0.1, 64] + 64 NeRF++-guided ones (ZT:1762-1799)
                m_idx = infinity_bkgr[k].flatten().nonzero().squeeze(1)
                s_m, d_m = start[m_idx].contiguous(), dk[m_idx].contiguous()
                Rm = s_m.shape[0]
                z = torch.linspace(0.1, 64.0, 192, device=dev).unsqueeze(0).expand(Rm, 192).contiguous()
                p = eng.segment_points(s_m, d_m, z)
                dists = z[:, 1:] - z[:, :-1]
                dists = torch.cat([dists, dists[:, -1:]], -1).contiguous()
                alpha = eng.nerf_alpha(w1.nerf, p.reshape(-1, 3), d_m.repeat_interleave(192, 0), dists.reshape(-1),
                                       w1.planes).reshape(Rm, 192)
                z = eng.importance_merge(z, alpha, 64)
                pts[m_idx] = eng.segment_points(s_m, d_m, z)
                if Z is not None:
                    Z[m_idx] = z
            if rec is not None:
                rec.setdefault("segments", []).append(dict(Z=Z, h_idx=h_idx, m_idx=m_idx, start=start, dir=dk))
            pathes.append(pts)
        # render_core reuses the index lists of the rays that continue (no second nonzero / host sync per segment) when it
        # is handed these very mask tensors (the cache holds them, so identity is a safe test)
        self._trace_cache = (list(converges), list(conv_idxs))
        return pathes, converges, directions, ior_ratios, infinity_bkgr, gradient_mesh, tirs[0]

    # ------------------------------------------------------------------ gradient of IORs_pred
    def _ior_torch(self, x):
        """IoRNetwork.forward (field.py:1046-1065) as a differentiable fp32 torch expression on the [M,3] hit points
        (M <= rays per bounce: a few thousand rows; the trace itself evaluates it on the tensor cores)."""
        pe = [x]
        for k in range(6):
            pe += [torch.sin(x * (2.0 ** k)), torch.cos(x * (2.0 ** k))]
        h = torch.cat(pe, -1)
        seq = self.IORs_pred.module0
        for i, act in ((0, True), (2, True), (4, False), (5, False)):
            h = F.linear(h, seq[i].effective_weight(), seq[i].bias)
            h = F.relu(h) if act else h
        return torch.sigmoid(h)

    def _replay_geometry(self, rays_o, rays_d, rec, pathes, directions, gradient_mesh, straight_through=True):
        """ray_trace (ZT:1571-1720) again as a differentiable function of IORs_pred, with every discrete decision (hit
        triangle, pass / total-internal-reflection mask, sample parameters z -- no_grad in the reference too) taken from
        the recorded trace: Moeller-Trumbore on the recorded triangle (DiffRender.py:61-124), interpolated vertex
        normal, IoR network, Snell.  Values are the trace's own (straight-through: v_trace + (v - v.detach())), so the
        forward outputs do not change; only the graph is added.  Returns (pathes, directions, gradient_mesh).
        straight_through=False returns the replayed values themselves (tests: they must reproduce the trace)."""
        if straight_through:
            st = lambda kernel, replay: replay if kernel is None else _StraightThrough.apply(kernel.detach(), replay)
        else:
            st = lambda kernel, replay: replay
        sc = self.scene
        # kernels: the two reverse kernels of csrc/bvh.cu + the path-point reduction; torch: the same chain as plain
        # differentiable torch expressions (CPU tests, cfg['replay_impl'] = 'torch' -- the two are tested against each other)
        kernels = straight_through and rays_o.is_cuda and self.cfg.get("replay_impl", "kernels") == "kernels"
        verts, faces, vnorm = sc.vertices.float(), sc.faces, sc.normals.float()
        start_r, dir_r = rays_o.float(), rays_d.float()
        starts, dirs, xs, gm = [start_r], [dir_r], [], []
        for i, b in enumerate(rec["bounces"]):
            hit_idx, ok_idx = b["hit_idx"], b["ok_idx"]
            o_c, d_c = start_r.index_select(0, hit_idx), dir_r.index_select(0, hit_idx)
            if kernels:
                x, normal = _HitFn.apply(o_c, d_c, (sc, b["tri"].to(torch.int32).contiguous(), b["inside"], b.get("x"),
                                                    b.get("n")))
            else:
                f = faces[b["tri"]]
                tv, tn = verts[f], vnorm[f]
                v0, e1, e2 = tv[:, 0], tv[:, 1] - tv[:, 0], tv[:, 2] - tv[:, 0]
                pvec = torch.cross(d_c, e2, dim=-1)
                inv_det = 1.0 / (e1 * pvec).sum(-1)
                tvec = o_c - v0
                u = (tvec * pvec).sum(-1) * inv_det
                qvec = torch.cross(tvec, e1, dim=-1)
                v = (d_c * qvec).sum(-1) * inv_det
                t = (e2 * qvec).sum(-1) * inv_det
                x = st(b.get("x"), o_c + t[:, None] * d_c)
                n = (1 - u - v)[:, None] * tn[:, 0] + u[:, None] * tn[:, 1] + v[:, None] * tn[:, 2]
                n = F.normalize(n / n.norm(dim=1, keepdim=True), dim=-1)
                normal = st(b.get("n"), -n if b["inside"] else n)
            xs.append(x)
            if ok_idx.numel() == 0:
                break
            gm.append(normal.index_select(0, ok_idx))
            if i + 1 >= len(rec["segments"]):
                break                                   # the ray leaving the third hit is not rendered
            eta = 1.0 / (self._ior_torch(x).reshape(-1, 1) * 1.0 + 1.0)
            eta = st(b["eta"].reshape(-1, 1) if b.get("eta") is not None else None, eta)
            if b["inside"]:
                eta = 1.0 / eta
            if kernels:
                s_next, d_next = _RefractFn.apply(x.index_select(0, ok_idx), normal.index_select(0, ok_idx),
                                                  d_c.index_select(0, ok_idx), eta.index_select(0, ok_idx).reshape(-1),
                                                  (starts_k(rec, i + 1, "start"), starts_k(rec, i + 1, "dir")))
                start_r, dir_r = s_next, d_next
            else:
                cos_i = (normal * -d_c).sum(-1, keepdim=True)
                sin2 = 1.0 - cos_i * cos_i
                eta, cos_k, n_k = eta[ok_idx], cos_i[ok_idx], normal[ok_idx]
                sin_t2 = sin2[ok_idx] * eta * eta
                d_tmp = eta * d_c[ok_idx] + (eta * cos_k - torch.sqrt(1.0 - sin_t2)) * n_k
                s_next = x[ok_idx] + d_tmp * 1e-5
                d_next = d_tmp / (torch.linalg.norm(d_tmp, dim=-1, keepdim=True) + 0.0001)
                start_r, dir_r = st(starts_k(rec, i + 1, "start"), s_next), st(starts_k(rec, i + 1, "dir"), d_next)
            starts.append(start_r)
            dirs.append(dir_r)
        new_pathes = []
        for k, sg in enumerate(rec["segments"]):
            s_k, d_k = starts[k], dirs[k]
            if not (s_k.requires_grad or d_k.requires_grad):
                new_pathes.append(pathes[k])
                continue
            end = s_k + d_k * 4.5
            if sg["h_idx"].numel() > 0:
                end = end.index_put((sg["h_idx"],), xs[k])
            delta = end - s_k
            if sg["m_idx"] is not None:
                delta = delta.index_put((sg["m_idx"],), d_k.index_select(0, sg["m_idx"]))
            if kernels:
                new_pathes.append(_SegPointsFn.apply(s_k, delta, sg["Z"], pathes[k]))
            else:
                pts = s_k[:, None, :] + delta[:, None, :] * sg["Z"][:, :, None]
                new_pathes.append(st(pathes[k], pts))
        if not straight_through:
            new_pathes = [p_ if p_.requires_grad else None for p_ in new_pathes]
        new_dirs = [dirs[k] if k < len(dirs) else directions[k] for k in range(len(directions))]
        new_gm = [gm[k] if k < len(gm) else gradient_mesh[k] for k in range(len(gradient_mesh))]
        if self.cfg.get("debug_geometry_grads"):
            for t_ in new_pathes + new_dirs + new_gm:
                if t_.requires_grad:
                    t_.retain_grad()
            self._geo_debug = dict(pathes=new_pathes, directions=new_dirs, gradient_mesh=new_gm)
        return new_pathes, new_dirs, new_gm

    def _eval_inner(self, pack, inv_s, pts, dirs, dists, params, step, is_train):
        """Inner SDF field + inner shading on the compact inner samples of segment 1 -> (alpha, colour, eikonal term)."""
        return _InnerFn.apply(pack, inv_s, pts, dirs, dists, *params)

    @staticmethod
    def _hit_from_inside(i):
        """The `inner` flag of the surface shader at the hits of segment i (ZT:1915: i % 2 != 0)."""
        return i % 2 != 0

    def _train_ior(self):
        return torch.is_grad_enabled() and not self.cfg.get("frozen_ior", False) and \
            any(p.requires_grad for p in self.IORs_pred.parameters())

    # ------------------------------------------------------------------ ZT:1835-2011
    def render_core(self, rays_o, rays_d, pathes, converges, directions, infinity_bkgr, gradient_mesh, ior_ratios,
                    human_poses=None, cos_anneal_ratio=0.0, step=None, is_train=True, is_nerf=False, prepared=None):
        eng = _engine()
        w1, wi, _ = prepared if prepared is not None else self._prepare()
        dev = rays_o.device
        R = converges[0].shape[0]
        T = torch.ones(R, 3, device=dev)
        normals_out = torch.zeros(R, 3, device=dev)
        spec_color_out, spec_light_out, spec_ref_out = (torch.zeros(R, 3, device=dev) for _ in range(3))
        colors, tmp, conv_idxs = [], {}, []
        exp_max1 = self.stage1_network.color_network.cfg["light_exp_max"]
        exp_maxi = self.color_network_inner.cfg["light_exp_max"]
        p1, p_in = _bank_params(w1), _bank_params(wi)
        freeze = self.cfg["freeze_inv_s_step"]
        frozen = freeze is not None and step is not None and step < freeze
        # geometry tensors that carry a graph (IORs_pred being trained, see _replay_geometry) stay attached
        geo_grad = torch.is_grad_enabled() and any(t_.requires_grad for t_ in list(pathes) + list(directions))
        keep = (lambda t_: t_) if geo_grad else (lambda t_: t_.detach())
        for i in range(len(pathes)):
            cand = keep(pathes[i])
            N, S = cand.shape[0], cand.shape[1] - 1
            cache = getattr(self, "_trace_cache", None)
            if cache is not None and i < len(cache[0]) and converges[i] is cache[0][i]:
                conv_idx = cache[1][i]                                    # from ray_trace: no host sync
            else:
                conv_idx = converges[i].flatten().nonzero().squeeze(1)    # foreign lists: one host sync per segment
            conv_idxs.append(conv_idx)
            dirs_i = keep(directions[i])
            if cand.requires_grad:
                # the path carries the IoR network's graph: points / interval lengths as differentiable torch expressions
                # (ZT:1853-1858), the inside-the-unit-sphere mask from their values
                pts = cand[:, :-1, :]
                dists = torch.linalg.norm(pts[:, 1:] - pts[:, :-1], dim=-1)
                dists = torch.cat([dists, dists[:, -1:]], -1)
                inner = torch.norm(pts.detach(), dim=-1) <= 1.0
            else:
                with torch.no_grad():
                    pts, dists, inner = eng.segment_geometry(cand)
            with torch.no_grad():
                inner_f = inner.reshape(-1)
                outer_idx = (~inner_f).nonzero().squeeze(1)
                inner_idx = inner_f.nonzero().squeeze(1) if i == 1 else None
                ray_of = lambda idx: torch.div(idx, S, rounding_mode="floor")      # sample -> ray (view dir)
            pts_f, dists_f = pts.reshape(-1, 3), dists.reshape(-1)
            alpha = torch.zeros(N * S, device=dev)
            color = torch.zeros(N * S, 3, device=dev)
            if outer_idx.numel() > 0:
                # NeRF++ of the STAGE-1 network on the samples outside the unit sphere (ZT:1876-1880)
                # (index_select: its backward is an atomic index_add, not the sort-based scatter of tensor[idx])
                a_o, c_o = _NerfFn.apply(w1, pts_f.index_select(0, outer_idx), dirs_i.index_select(0, ray_of(outer_idx)),
                                         dists_f.index_select(0, outer_idx), *p1)
                alpha = alpha.index_put((outer_idx,), a_o)
                color = color.index_put((outer_idx,), c_o)
            if i == 1 and inner_idx.numel() > 0:
                # inner SDF field + inner shading on segment 1 (ZT:1883-1906)
                inv_s = torch.exp(self.deviation_network_inner.variance * 10.0)
                a_i, c_i, gerr = self._eval_inner((wi, float(cos_anneal_ratio), exp_maxi, not frozen), inv_s,
                                                  pts_f.index_select(0, inner_idx),
                                                  dirs_i.index_select(0, ray_of(inner_idx)),
                                                  dists_f.index_select(0, inner_idx), p_in, step, is_train)
                alpha = alpha.index_put((inner_idx,), a_i)
                color = color.index_put((inner_idx,), c_i)
                inv_s_c = inv_s.clip(1e-6, 1e6)
                tmp["std"] = torch.mean(1.0 / (inv_s_c.detach() if frozen else inv_s_c))
                tmp["gradient_error"] = gerr
            alpha, color = alpha.view(N, S), color.view(N, S, 3)
            # linear-space compositing (ZT:1942-1951): one kernel forward, one backward (transmittance recomputed)
            rgb_lin, t_end = eng.SegCompositeFn.apply(alpha, color)
            color_now = rgb_lin * T
            T = T * t_end[:, None]
            n_hit = conv_idx.numel()
            if n_hit > 0:
                p_hit = cand[:, -1, :].index_select(0, conv_idx)
                holder = {}
                c_s, trans, nov = _SurfaceFn.apply((w1, exp_max1, holder), p_hit, keep(gradient_mesh[i]).contiguous(),
                                                   dirs_i.index_select(0, conv_idx), *p1)
                if self._hit_from_inside(i):
                    c_s = torch.zeros_like(c_s)                    # inside the object: field.py:969
                tn = torch.clamp(1.0 - nov[:, None], 0.0, 1.0)
                rw = torch.clamp(0.04 + 0.96 * tn * tn * tn * tn * tn, 0.0, 1.0)
                color_now = color_now.index_add(0, conv_idx, srgb_to_linear(c_s) * T.index_select(0, conv_idx))
                if i == 0 and not is_train:
                    with torch.no_grad():
                        ex = eng.surface_extras(w1, holder["tape"], exp_max1)
                        normals_out[conv_idx] = (F.normalize(gradient_mesh[i].reshape(-1, 3), dim=-1) + 1.0) * 0.5
                        spec_color_out[conv_idx], spec_light_out[conv_idx], spec_ref_out[conv_idx] = \
                            ex["specular_color"], ex["specular_light"], ex["specular_ref"]
                T = T.index_select(0, conv_idx) * ((1.0 - rw) * trans[:, None])        # refraction_coefficient, ZT:1966
                colors.append(color_now)
            else:
                colors.append(color_now)
                break
        total = colors[-1]
        for i in range(len(colors) - 1, 0, -1):
            total = colors[i - 1].index_add(0, conv_idxs[i - 1], total)
        ray_rgb = torch.clamp(linear_to_srgb(total), min=0.0, max=1.0)
        return {
            "ray_rgb": ray_rgb,
            "gradient_error": tmp.get("gradient_error", torch.zeros(1, device=dev)),
            "acc": torch.ones(1, device=dev),
            "normal": normals_out, "specular_color": spec_color_out, "specular_light": spec_light_out,
            "specular_ref": spec_ref_out,
            "std": tmp.get("std", torch.zeros(1, device=dev)),
        }

    # ------------------------------------------------------------------ ZT:1442-1466
    def render(self, rays_o, rays_d, near=None, far=None, human_poses=None, perturb_overwrite=-1, cos_anneal_ratio=0.0,
               is_train=True, step=None, is_nerf=False):
        prepared = self._prepare()
        rec = {} if (is_train and self._train_ior()) else None
        pathes, converges, directions, ior_ratios, infinity_bkgr, gradient_mesh, tir_mask = \
            self.ray_trace(rays_o, rays_d, prepared=prepared, rec=rec)
        if rec is not None and rec.get("bounces"):
            # IORs_pred is trained through the path geometry (ZT:1642-1684): rebuild it with a graph
            pathes, directions, gradient_mesh = self._replay_geometry(rays_o, rays_d, rec, pathes, directions, gradient_mesh)
        ret = self.render_core(rays_o, rays_d, pathes, converges, directions, infinity_bkgr, gradient_mesh, ior_ratios,
                               human_poses, cos_anneal_ratio=cos_anneal_ratio, step=step, is_train=is_train,
                               is_nerf=is_nerf, prepared=prepared)
        ret["tir_mask"] = tir_mask
        return ret

    # ------------------------------------------------------------------ ZT:1090-1123
    @torch.no_grad()
    def nvs(self, pose, K, h, w, chunk=4096):
        """Novel-view synthesis through the refraction path: `pose` [3,4] world-to-camera, `K` [3,3] -> ray_rgb [h,w,3]
        numpy (eval settings of the reference: is_train=False, step 300000)."""
        from . import feeder
        dev = self.deviation_network_inner.variance.device
        as_t = lambda a: (torch.from_numpy(np.asarray(a, dtype=np.float32)) if not torch.is_tensor(a) else a.float()).to(dev)
        K_, pose_ = as_t(K).unsqueeze(0), as_t(pose).unsqueeze(0)
        batch, rn, _, _ = feeder.construct_ray_batch(torch.zeros(1, 3, h, w, device=dev), K_)
        colors = []
        for ri in range(0, rn, chunk):
            rays_o, rays_d = feeder.world_rays(batch["dirs"][ri:ri + chunk], batch["idxs"][ri:ri + chunk], pose_)
            out = self.render(rays_o.contiguous(), rays_d.contiguous(), None, None, None, 0, 0, is_train=False, step=300000)
            colors.append(out["ray_rgb"])
        return torch.cat(colors, 0).reshape(h, w, 3).cpu().numpy()

    # ------------------------------------------------------------------ ZT:1209-1257
    TEST_KEYS = ("ray_rgb", "gradient_error", "normal", "tir_mask", "specular_light", "specular_color", "specular_ref")

    def set_eval_source(self, fn):
        """fn(index) -> {'rays_o', 'rays_d', 'rgbs' [h*w,3], 'h', 'w'[, 'gt_depth', 'gt_mask']} (see
        NeROShapeRenderer.set_eval_source / feeder.image_eval_source)."""
        self.eval_source = fn

    def test_step(self, index, step):
        """One test view through the refraction path in chunks of cfg['test_ray_num'] rays; ray_rgb / gt_rgb are masked by
        the total-internal-reflection mask like the reference's (ZT:1247-1249)."""
        if getattr(self, "eval_source", None) is None:
            raise RuntimeError("no eval source attached: call set_eval_source(fn) (the image database is outside the hot path)")
        src = self.eval_source(index)
        rays_o, rays_d = src["rays_o"].float(), F.normalize(src["rays_d"].float(), dim=-1)
        h, w = int(src["h"]), int(src["w"])
        rn, trn = rays_o.shape[0], self.cfg["test_ray_num"]
        outputs = {k: [] for k in self.TEST_KEYS}
        with torch.no_grad():
            for ri in range(0, rn, trn):
                cur = self.render(rays_o[ri:ri + trn].contiguous(), rays_d[ri:ri + trn].contiguous(), None, None, None, 0, 0,
                                  is_train=False, step=step, is_nerf=self.is_nerf)
                for k in self.TEST_KEYS:
                    outputs[k].append(cur[k].detach())
        outputs = {k: torch.cat(v, 0) for k, v in outputs.items()}
        tm = outputs["tir_mask"]
        outputs["loss_rgb"] = self.compute_rgb_loss(outputs["ray_rgb"] * tm, src["rgbs"] * tm)
        outputs["gt_rgb"] = (src["rgbs"] * tm).reshape(h, w, 3)
        outputs["ray_rgb"] = (outputs["ray_rgb"] * tm).reshape(h, w, 3)
        for k in ("gt_depth", "gt_mask"):
            if k in src:
                outputs[k] = torch.as_tensor(src[k]).unsqueeze(-1)
        self.zero_grad()
        return outputs

    def forward(self, data):
        step = data["step"]
        if "eval" in data:
            return self.test_step(data["index"], step)
        if self.ray_source is None:
            raise NotImplementedError("dataset ingest is outside the hot path: attach a ray source and call render()")
        batch = self.ray_source(step, self.cfg["train_ray_num"])
        rays_d = F.normalize(batch["rays_d"], dim=-1)
        out = self.render(batch["rays_o"], rays_d, None, None, None, -1, self.get_anneal_val(step), is_train=True,
                          step=step, is_nerf=self.is_nerf)
        tm = out["tir_mask"].detach()
        out["loss_rgb"] = self.compute_rgb_loss(out["ray_rgb"] * tm, batch["rgbs"] * tm)        # ZT:1272
        return out

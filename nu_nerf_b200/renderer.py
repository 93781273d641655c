"""Drop-in boundary for the reference's network/renderer.py ("NZ"): the renderers of the NON-zero-thickness pipeline
(SURVEY 8f row 1), i.e. the refraction path through a thin glass shell of learned thickness around the outer mesh.

    - from network.renderer import name2renderer
    + from nu_nerf_b200.renderer import name2renderer          # run_training.py:17-20 picks this module when
                                                               # cfg['zero_thickness'] is False

What differs from the zero-thickness classes (renderer_zerothick.py / renderer_stage2.py) and is built here:
  stage 1 (NeROShapeRenderer, NZ:102-905)
  * render_core (NZ:738-859): `loss_normal` = sum_j w_j max(grad_j . dir_j, 0) -- a second pass of the compositing kernels
    with that scalar as a colour channel, forward and backward (its d alpha joins the main one, its d colour reaches the SDF
    gradient and from there the reverse-over-reverse passes); `color_spec` / `color_bkgr` for the rays whose 65th sample
    lies in the unit sphere only (NZ:798-821), with `sphere_direction` the probe rows are [IDE(d) | IDE(exit direction)].
  * train_step adds `loss_mask` = l1(masks, acc) for synthetic data (NZ:478).
  stage 2 (Stage2Renderer, NZ:907-2378)
  * ray_trace (NZ:1610-2148): per bounce the curvature-radius shell offset with two refractions (nu_nerf_b200/shell.py,
    pinned bounce by bounce to the reference), the interpolated Gaussian curvature of the hit (Scene.Dintersect()['g_k'],
    tracer.discrete_gaussian_curvature -- the stated replacement of the PyMesh attribute, parity against PyMesh unpinned),
    IoR = 1 / (IORs_pred + 0.6), ThicknessNetwork * 0.01, retro-active un-convergence of rays that miss the mesh from the
    inside (NZ:1662-1672), 64 / 128 (64 + 2 x 32 SDF-guided) / 64 samples per segment, 64 inverse-depth samples for rays
    that leave the scene (NZ:2140-2143).  Kernels: BVH closest hit + re-intersection, IoR / thickness MLPs, inner-SDF
    queries + up-sampling rounds, path points, and the closed-form bounce itself (csrc/shell.cu: one launch per bounce
    forward, one backward -- the hand-derived adjoint, so that the loss reaches IORs_pred AND thickness_pred through it;
    shell.shell_bounce, the torch restatement, is the cross-check: cfg['shell_impl'] = 'torch').
  * the inner field's shader is AppShadingNetwork_SpecInner (field.py:1320: PE-8 positions, PE-2 refraction inputs,
    refraction light clamped at exp(-0.2)) -- shade_encode_*_var_kernel<8, 2, .>, exp_max_refrac of the mixing kernels.
  * render_core (NZ:2155-2353): surface shading flagged `inner` for every segment but the first (NZ:2244), `loss_occ` key.
  * render(rays_o, rays_d, mask, near, far, ...) (NZ:1482) and the masked losses of train_step / test_step (NZ:1297-1299,
    :1364).
  both: shader_config.sphere_direction (field.py:594-597, :641-651, :675-680; every zero_thickness: False config sets it).
  * the occlusion-probe loss of the inner field (NZ:2222-2230, :1580-1608) on the stage-1 probe machinery.
Not built (raises NotImplementedError): the human_light shader variant (no config of the reference sets it).
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .field import (SDFNetwork, SingleVarianceNetwork, NeRFNetwork, AppShadingNetwork_SpecInner, AppShadingNetwork_S2,
                    IoRNetwork, ThicknessNetwork)
from .renderer_stage2 import Stage2Renderer as _ZTStage2Renderer, _HitFn, _engine, _zf
from .renderer_zerothick import NeROShapeRenderer as _ZTShapeRenderer, load_cfg, linear_to_srgb, _SdfValueFn
from .shell import shell_bounce, signed_normal, outside_depths


class _ShellFn(torch.autograd.Function):
    """The non-zero-thickness bounce (NZ:1690-2009) on the M hit rays as ONE launch forward (`shell_bounce_fwd_kernel`) and
    one backward (`shell_bounce_bwd_kernel`, the hand-derived adjoint; csrc/shell.cu, pointwise.cuh).  The torch
    restatement nu_nerf_b200/shell.shell_bounce is its cross-check (cfg['shell_impl'] = 'torch')."""

    @staticmethod
    def forward(ctx, x, n, d, gk, ior_sig, th_sig, inside):
        eng = _engine()
        c = lambda t: t.detach().contiguous().float()
        x, n, d, gk, ior_sig, th_sig = c(x), c(n), c(d), c(gk), c(ior_sig), c(th_sig)
        M, dev = x.shape[0], x.device
        ok, tir = (torch.empty(M, dtype=torch.uint8, device=dev) for _ in range(2))
        x_mod, start, dirn = (torch.empty(M, 3, device=dev) for _ in range(3))
        ratio = torch.empty(M, device=dev)
        eng.call("nunerf_shell_bounce", x.data_ptr(), n.data_ptr(), d.data_ptr(), gk.data_ptr(), ior_sig.data_ptr(),
                 th_sig.data_ptr(), M, int(inside), ok.data_ptr(), tir.data_ptr(), x_mod.data_ptr(), start.data_ptr(),
                 dirn.data_ptr(), ratio.data_ptr())
        ctx.save_for_backward(x, n, d, gk, ior_sig, th_sig, ok)
        ctx.inside = bool(inside)
        ctx.mark_non_differentiable(ok, tir)
        return ok, tir, x_mod, start, dirn, ratio

    @staticmethod
    def backward(ctx, _g_ok, _g_tir, g_xmod, g_start, g_dir, g_ratio):
        eng = _engine()
        x, n, d, gk, ior_sig, th_sig, ok = ctx.saved_tensors
        M, dev = x.shape[0], x.device
        z = lambda g, *shape: torch.zeros(*shape, device=dev) if g is None else g.contiguous().float()
        g_xmod, g_start, g_dir, g_ratio = z(g_xmod, M, 3), z(g_start, M, 3), z(g_dir, M, 3), z(g_ratio, M)
        d_x, d_n, d_d = (torch.empty(M, 3, device=dev) for _ in range(3))
        d_gk, d_ior, d_th = (torch.empty(M, device=dev) for _ in range(3))
        eng.call("nunerf_shell_bounce_bwd", x.data_ptr(), n.data_ptr(), d.data_ptr(), gk.data_ptr(), ior_sig.data_ptr(),
                 th_sig.data_ptr(), ok.data_ptr(), M, int(ctx.inside), g_start.data_ptr(), g_dir.data_ptr(),
                 g_ratio.data_ptr(), g_xmod.data_ptr(), d_x.data_ptr(), d_n.data_ptr(), d_d.data_ptr(), d_gk.data_ptr(),
                 d_ior.data_ptr(), d_th.data_ptr())
        return d_x, d_n, d_d, d_gk.reshape(gk.shape), d_ior.reshape(ior_sig.shape), d_th.reshape(th_sig.shape), None


def shell_bounce_kernels(x, normal, d, g_k, ior_sig, thick_sig, inside):
    """shell.shell_bounce with the bounce on the device kernels: same arguments, same result dict."""
    ok, tir, x_mod, start, dirn, ratio = _ShellFn.apply(x, normal, d, g_k.reshape(-1), ior_sig.reshape(-1),
                                                        thick_sig.reshape(-1), inside)
    ok = ok.bool()
    ok_idx = ok.nonzero().squeeze(1)
    return {"ok": ok, "ok_idx": ok_idx, "tir": tir.bool(), "x_mod": x_mod, "normal": normal.index_select(0, ok_idx),
            "ratio": ratio.index_select(0, ok_idx).reshape(-1, 1), "start": start.index_select(0, ok_idx),
            "dir": dirn.index_select(0, ok_idx)}


class _InnerOccFn(torch.autograd.Function):
    """renderer_stage2._InnerFn with the outputs the inner-field occlusion loss reads (NZ:2222-2230): the predicted
    occlusion probability (differentiable) and, without a graph, the samples' points, SDF values, SDF gradients, ray
    directions and reflected directions."""

    @staticmethod
    def forward(ctx, pack, inv_s, pts, dirs, dists, *params):
        eng = _engine()
        w, cos_anneal, exp_max, want_inv_s = pack
        t = eng.inner_tape(pts, dirs, dists, cos_anneal, exp_max)
        eng.inner_forward(w, t)
        ctx.w, ctx.tape, ctx.n, ctx.want_inv_s = w, t, len(params), want_inv_s
        aux = (t.pts_in, t.sdf.sdf[:, 0], t.sdf.grad, t.dirs_in, t.refl)
        ctx.mark_non_differentiable(*aux)
        return (t.a_in, t.c_in, t.gerr, t.occ[:, None]) + aux

    @staticmethod
    def backward(ctx, d_alpha, d_color, d_gerr, d_occ, *_unused):
        eng = _engine()
        w, t = ctx.w, ctx.tape
        dev, M = t.pts_in.device, t.n_in
        geo = any(ctx.needs_input_grad[2:5])
        w.bank.zero_grads()
        g = eng.inner_backward(w, t, _zf(d_alpha, dev, M), _zf(d_color, dev, M, 3), _zf(d_gerr, dev, M), None, None,
                               ctx.want_inv_s, d_occ=None if d_occ is None else d_occ.contiguous().float().reshape(-1),
                               want_geo=geo)
        w.bank.backward()
        ctx.tape = None
        d_inv = g["inv_s"].reshape(()) if ctx.want_inv_s and "inv_s" in g else None
        return (None, d_inv, g.get("d_pts"), g.get("d_dirs"), g.get("d_dists")) + (None,) * ctx.n


class _InnerOccAdapter:
    """compute_occ_loss / occ_probability of the stage-1 class (ZT:695-723 == NZ:1580-1608, field.py:501-554) bound to the
    INNER field of stage 2: `sdf_inter_fun` = sdf_network_inner.sdf, `deviation_network_inner` (NZ:1025, :1596)."""
    compute_occ_loss = _ZTShapeRenderer.compute_occ_loss
    occ_probability = _ZTShapeRenderer.occ_probability

    def __init__(self, r):
        self.cfg, self.deviation_network = r.cfg, r.deviation_network_inner

    def _prepare(self):
        raise RuntimeError("the inner-field occlusion loss is always called with prepared operands")


class _RenderCoreNZFn(torch.autograd.Function):
    """NZ:738-859 as one autograd node: engine.core_forward(nz=True) / core_backward(d_normal=...).  Outputs of the
    zero-thickness node plus loss_normal [R,1] (differentiable) and the candidate-ray mask of the specular probe."""

    @staticmethod
    def forward(ctx, pack, inv_s, *params):
        eng = _engine()
        w, rays_o, rays_d, z_vals, cos_anneal, is_nerf, exp_max, want_inv_s, want_weights = pack
        t, rgb, acc, bkgr, wts = eng.core_forward(w, rays_o, rays_d, z_vals, cos_anneal, is_nerf, exp_max,
                                                  want_weights=want_weights, nz=True)
        ctx.tape, ctx.w, ctx.n_params, ctx.want_inv_s = t, w, len(params), want_inv_s
        dev = rgb.device
        if t.n_in > 0:
            gerr, trans, met, occ = t.gerr, t.trans[:, None], t.metallic[:, None], t.occ[:, None]
            aux = (t.pts_in, t.sdf.sdf[:, 0], t.sdf.grad, t.dirs_in, t.refl)
        else:
            gerr, trans, met, occ = (torch.zeros(1, device=dev), torch.zeros(0, 1, device=dev),
                                     torch.zeros(0, 1, device=dev), torch.zeros(0, 1, device=dev))
            aux = tuple(torch.zeros(0, 3, device=dev) if i != 1 else torch.zeros(0, device=dev) for i in range(5))
        spec = t.ls_.head[:, :3].clone()
        ctx.mark_non_differentiable(wts, t.cand, *aux)
        return (rgb, acc, bkgr, gerr, trans, met, spec, occ, t.loss_normal, wts, t.cand) + aux

    @staticmethod
    def backward(ctx, d_rgb, d_acc, d_bkgr, d_gerr, d_trans, d_met, d_spec, d_occ, d_ln, *_unused):
        eng = _engine()
        t = ctx.tape
        c = lambda x: None if x is None else x.contiguous().float()
        has = t.n_in > 0
        ctx.w.bank.zero_grads()
        g = eng.core_backward(ctx.w, t, c(d_rgb), c(d_acc), c(d_bkgr), c(d_gerr) if has else None,
                              c(d_trans).reshape(-1) if (has and d_trans is not None) else None,
                              c(d_met).reshape(-1) if (has and d_met is not None) else None, c(d_spec), ctx.want_inv_s,
                              d_occ=c(d_occ).reshape(-1) if (has and d_occ is not None) else None, d_normal=c(d_ln))
        ctx.w.bank.backward()
        ctx.tape = None
        return (None, g.get("inv_s").reshape(()) if ctx.want_inv_s and "inv_s" in g else None,
                *([None] * ctx.n_params))


class NeROShapeRenderer(_ZTShapeRenderer):
    """network/renderer.py:102-905: the stage-1 renderer of the non-zero-thickness pipeline.  Same modules, sampling and
    field evaluation as the zero-thickness class; render_core additionally returns `loss_normal` (NZ:766-780: the
    composited back-facing part of the SDF gradient, for NormalOrientationLoss, network/loss.py:105-112), evaluates the
    specular probe -- with `sphere_direction` on [IDE(d) | IDE(exit direction)] -- and reports `color_spec` / `color_bkgr`
    only for the rays whose 65th sample lies in the unit sphere (NZ:798-821); train_step adds `loss_mask` for synthetic
    data (NZ:478)."""
    default_cfg = {**_ZTShapeRenderer.default_cfg, "train_ray_num": 1024, "downsample_ratio": 1.0, "get_mask": False}

    def __init__(self, cfg, training=True):
        super().__init__(cfg, training=training)
        self.get_mask = self.cfg["get_mask"]

    def render_core(self, rays_o, rays_d, z_vals, human_poses=None, cos_anneal_ratio=0.0, step=None, is_train=True,
                    is_nerf=False, prepared=None, occ_perm=None):
        w = prepared if prepared is not None else self._prepare()
        params = [p for d in w.bank.denses if d.has_grad for p in (d.v, d.g, d.bias) if p is not None]
        params = list({id(p): p for p in params}.values())
        freeze = self.cfg["freeze_inv_s_step"]
        frozen = freeze is not None and step is not None and step < freeze
        inv_s = torch.exp(self.deviation_network.variance * 10.0)
        exp_max = self.color_network.cfg["light_exp_max"]
        pack = (w, rays_o, rays_d, z_vals, float(cos_anneal_ratio), bool(is_nerf), exp_max, not frozen, not is_train)
        (rgb, acc, bkgr, gerr, trans, met, spec, occ, loss_normal, weights, cand, pts_in, sdf_in, grad_in, dirs_in,
         refl_in) = _RenderCoreNZFn.apply(pack, inv_s, *params)
        inv_s_c = inv_s.clip(1e-6, 1e6)
        if frozen:
            inv_s_c = inv_s_c.detach()
        has_inner = trans.shape[0] > 0
        outputs = {
            "ray_rgb": rgb, "gradient_error": gerr, "loss_normal": loss_normal, "acc": acc,
            "color_bkgr": bkgr[cand],                                           # NZ:812 (boolean mask: one host sync)
            "color_spec": linear_to_srgb(torch.exp(torch.clamp(spec, max=exp_max)))[cand],
            "std": torch.mean(1.0 / inv_s_c) if has_inner else torch.zeros(1, device=rgb.device),
        }
        if step is not None and step < 1000:
            zf = z_vals.float()
            dist = torch.cat([zf[:, 1:] - zf[:, :-1], zf[:, -1:] - zf[:, -2:-1]], -1)
            pts_all = rays_o.float()[:, None, :] + rays_d.float()[:, None, :] * (zf + dist * 0.5)[..., None]
            mask = torch.norm(pts_all, dim=-1) < 1.2
            sdf_pts = pts_all[mask].contiguous()
            outputs["sdf_pts"] = sdf_pts
            outputs["sdf_vals"] = _SdfValueFn.apply((w, sdf_pts), *params) if sdf_pts.shape[0] > 0 else \
                torch.zeros(0, device=rgb.device)
        if self.cfg["apply_occ_loss"]:
            if has_inner and step is not None:
                outputs["loss_occ"] = self.compute_occ_loss({"occ_prob": occ, "reflective": refl_in}, pts_in, sdf_in,
                                                            grad_in, dirs_in, step, prepared=w, perm=occ_perm)
            else:
                outputs["loss_occ"] = torch.zeros(1, device=rgb.device)
        if has_inner:
            outputs["transmission"] = trans
            outputs["metallic"] = met
        if not is_train:
            outputs.update(self.compute_validation_info(z_vals, rays_o, rays_d, weights, human_poses, step, prepared=w))
        return outputs

    def train_step(self, step):
        outputs = super().train_step(step)
        if self.is_nerf:                                                        # NZ:476-478
            batch = self._last_batch
            if "masks" not in batch:
                raise KeyError("the non-zero-thickness stage-1 train_step needs batch['masks'] for synthetic data (NZ:478)")
            outputs["loss_mask"] = F.l1_loss(batch["masks"].reshape(outputs["acc"].shape).float(), outputs["acc"],
                                             reduction="mean")
        return outputs


class Stage2Renderer(_ZTStage2Renderer):
    """network/renderer.py:907-2378."""
    default_cfg = {**_ZTStage2Renderer.default_cfg, "downsample_ratio": 1.0, "get_mask": False}

    def __init__(self, cfg, training=True):
        nn.Module.__init__(self)
        from .tracer import Scene
        self.cfg = {**self.default_cfg, **cfg}
        self.is_nerf = self.cfg["is_nerf"]
        self.get_mask = self.cfg["get_mask"]
        if (self.cfg["sdf_n_layers"], self.cfg["sdf_freq"], self.cfg["sdf_d_out"]) != (8, 6, 257) or \
                not self.cfg["clip_sample_variance"]:
            raise NotImplementedError("the B200 engine is built for the 8x256 / PE-6 / 257-output SDF network")
        # construction order = the reference's (NZ:958-1019): same seed => bit-identical initial parameters
        self.nerf_network = NeRFNetwork(D=8, d_in=4, d_in_view=3, W=256, multires=10, multires_view=4, skips=(4,))
        self.IORs = nn.Parameter(torch.zeros(10))
        cfg1 = cfg["stage1_cfg_dir"]
        cfg1 = dict(cfg1) if isinstance(cfg1, dict) else load_cfg(cfg1)
        cfg1.setdefault("precision", self.cfg["precision"])
        self.stage1_network = NeROShapeRenderer(cfg1, training=False)
        ckpt = cfg["stage1_ckpt_dir"]
        ckpt = ckpt if isinstance(ckpt, dict) else torch.load(ckpt, map_location="cpu")
        self.stage1_network.load_state_dict(ckpt["network_state_dict"], strict=False)
        self.infinity_far_bkgr = self.stage1_network.infinity_far_bkgr        # NZ:992 (registered a second time)
        self._mesh = cfg["stage1_mesh_dir"]
        self.scene = None
        self._scene_cls = Scene
        self.IORs_pred = IoRNetwork()
        self.IoRint_pred = IoRNetwork()
        self.thickness_pred = ThicknessNetwork()
        self.color_network = AppShadingNetwork_S2(self.cfg["shader_config"], self.stage1_network)
        self.sdf_network_inner = SDFNetwork(d_out=257, d_in=3, d_hidden=256, n_layers=8, skip_in=(4,), multires=6,
                                            bias=self.cfg["sdf_bias"], scale=1.0,
                                            geometric_init=self.cfg["geometry_init"])
        self.deviation_network_inner = SingleVarianceNetwork(init_val=self.cfg["inv_s_init"],
                                                             activation=self.cfg["std_act"])
        self.color_network_inner = AppShadingNetwork_SpecInner(self.cfg["shader_config"])
        self.sdf_network_inner._query = self._sdf_inner_query
        self.ray_source = None

    # ------------------------------------------------------------------ engine operands
    def _prepare(self):
        prepared = super()._prepare()
        eng = _engine()
        key = (self._planes(),) + tuple(p.data_ptr() for p in self.thickness_pred.parameters())
        if getattr(self, "_w_thick", None) is None or self._w_thick_key != key:
            dev = self.deviation_network_inner.variance.device
            self._w_thick, self._w_thick_key = eng.IorWeights(self.thickness_pred, self._planes(), dev), key
        self._w_thick.refresh()
        return prepared

    @staticmethod
    def _hit_from_inside(i):
        return i != 0                                            # NZ:2244

    def _train_ior(self):
        return torch.is_grad_enabled() and not self.cfg.get("frozen_ior", False) and \
            any(p.requires_grad for m in (self.IORs_pred, self.thickness_pred) for p in m.parameters())

    def _curvature_with_graph(self, x, tri, g_k):
        """The interpolated Gaussian curvature of the hit as a function of the hit point (DiffRender.py:113-116: the
        barycentric weights are differentiable w.r.t. the ray, so in the reference the curvature radius of the shell moves
        with the refracted path).  x [M,3] lies on the plane of triangle tri [M]: its barycentric coordinates are affine in x.
        Value = the trace's own g_k (straight-through), gradient = that of the interpolation."""
        if not x.requires_grad:
            return g_k
        sc = self.scene
        f = sc.faces[tri]
        tv = sc.vertices.float()[f]                                            # [M,3,3]
        kf = sc.gaussian_curvatures.float()[f].squeeze(-1)                      # [M,3]
        e1, e2, w = tv[:, 1] - tv[:, 0], tv[:, 2] - tv[:, 0], x - tv[:, 0]
        d00, d01, d11 = (e1 * e1).sum(-1), (e1 * e2).sum(-1), (e2 * e2).sum(-1)
        d20, d21 = (w * e1).sum(-1), (w * e2).sum(-1)
        den = d00 * d11 - d01 * d01
        u, v = (d11 * d20 - d01 * d21) / den, (d00 * d21 - d01 * d20) / den
        gk = ((1.0 - u - v) * kf[:, 0] + u * kf[:, 1] + v * kf[:, 2]).reshape(-1, 1)
        return g_k.detach() + (gk - gk.detach())

    @staticmethod
    def _sigmoid_mlp(net, x):
        """IoRNetwork / ThicknessNetwork.forward (field.py:1046-1081) as a differentiable fp32 torch expression on the
        [M,3] hit points (M <= rays per bounce)."""
        pe = [x]
        for k in range(6):
            pe += [torch.sin(x * (2.0 ** k)), torch.cos(x * (2.0 ** k))]
        h = torch.cat(pe, -1)
        seq = net.module0
        for i, act in ((0, True), (2, True), (4, False), (5, False)):
            h = F.linear(h, seq[i].effective_weight(), seq[i].bias)
            h = F.relu(h) if act else h
        return torch.sigmoid(h)

    # ------------------------------------------------------------------ NZ:1610-2148
    def ray_trace(self, rays_o, rays_d, mask=None, prepared=None, trace=None):
        """-> (pathes, converges, directions, ior_ratios, infinity_bkgr, gradient_mesh, tir_mask) as the reference's.
        Under autograd with trainable IoR / thickness networks the returned geometry carries their graph (hit points
        through hit_interp_bwd_kernel, the two MLPs as fp32 torch expressions, the bounce through shell_bounce); every
        discrete decision (hit triangle, pass masks, sample parameters) is taken without one, as in the reference.
        trace (dict, tests): receives the hit records per bounce; a "z_1" entry replaces the up-sampled parameters of
        segment 1 (the CDF inversion amplifies MLP rounding, SURVEY 7.3: parity of what follows is checked on equal z)."""
        eng = _engine()
        w1, wi, wior = prepared if prepared is not None else self._prepare()
        wth = self._w_thick
        grad = self._train_ior()
        dev = rays_o.device
        o, d = rays_o.float().contiguous(), rays_d.float().contiguous()
        next_start, next_dir = o, d
        starts, directions = [o], [d]
        intersections, converges, infinity_bkgr, ior_ratios, gradient_mesh, tirs = [], [], [], [], [], []
        hit_idxs, conv_idxs = [], []
        inside = False
        for i in range(3):
            with torch.no_grad():
                info, hit = self.scene.Dintersect(next_start.detach().contiguous(), next_dir.detach().contiguous())
            if trace is not None:
                trace[f"trace_hit_{i}"], trace[f"trace_tri_{i}"] = hit.float(), info["faces_ind"]
            hit_idx = hit.nonzero().squeeze(1)
            if i == 1 and hit_idx.numel() < next_start.shape[0]:
                # rays that entered the object and find no way out (open / leaking mesh) are taken back: their first hit
                # no longer counts as converged and they leave every per-ray list of segment 1 (NZ:1662-1672).  The
                # reference forgets ior_ratios[0] there and fails later in render_core; it is filtered here as well.
                keep = hit_idx
                conv0 = torch.zeros_like(converges[0])
                conv0[conv_idxs[0][keep]] = True
                converges[0] = conv0
                conv_idxs[0] = conv_idxs[0][keep]
                next_start, next_dir = next_start[keep], next_dir[keep]
                gradient_mesh[0], ior_ratios[0] = gradient_mesh[0][keep], ior_ratios[0][keep]
                starts[1], directions[1] = next_start, next_dir
                info = {k_: v[keep] for k_, v in info.items()}
                hit = hit[keep]
                hit_idx = torch.arange(keep.numel(), device=dev)
            N = next_start.shape[0]
            converged = hit.reshape(-1, 1)
            infinity_bkgr.append(~converged)
            x_c = info["x"].index_select(0, hit_idx).contiguous()
            n_c = signed_normal(info["n"].index_select(0, hit_idx), inside).contiguous()
            g_k = info["g_k"].index_select(0, hit_idx).reshape(-1, 1)
            d_c = next_dir.index_select(0, hit_idx).contiguous()
            M = hit_idx.numel()
            if M > 0 and grad:
                tri = info["faces_ind"].index_select(0, hit_idx).to(torch.int32).contiguous()
                x_c, n_c = _HitFn.apply(next_start.index_select(0, hit_idx), d_c, (self.scene, tri, inside, x_c, n_c))
                g_k = self._curvature_with_graph(x_c, tri.long(), g_k)
                ior_sig = self._sigmoid_mlp(self.IORs_pred, x_c)
                th_sig = self._sigmoid_mlp(self.thickness_pred, x_c)
            elif M > 0:
                with torch.no_grad():
                    ior_sig = eng.ior_forward(wior, x_c).reshape(-1, 1)
                    th_sig = eng.ior_forward(wth, x_c).reshape(-1, 1)
            else:
                ior_sig = th_sig = torch.zeros(0, 1, device=dev)
            if M > 0 and x_c.is_cuda and self.cfg.get("shell_impl", "kernels") == "kernels":
                b = shell_bounce_kernels(x_c, n_c, d_c, g_k, ior_sig, th_sig, inside)        # one launch (csrc/shell.cu)
            else:
                b = shell_bounce(x_c, n_c, d_c, g_k, ior_sig, th_sig, inside)                # torch restatement (cross-check)
            ok_idx = b["ok_idx"]
            converged_out = torch.zeros(N, 1, dtype=torch.bool, device=dev)
            converged_out[hit_idx[ok_idx]] = True
            tir = torch.ones(N, 1, dtype=torch.bool, device=dev)
            tir[hit_idx] = b["tir"].reshape(-1, 1)
            tirs.append(tir)
            next_dir, next_start = b["dir"], b["start"]
            directions.append(next_dir)
            starts.append(next_start)
            converges.append(converged_out)
            intersections.append(b["x_mod"])
            hit_idxs.append(hit_idx)
            conv_idxs.append(hit_idx[ok_idx])
            if ok_idx.numel() == 0:
                break
            gradient_mesh.append(b["normal"])
            ior_ratios.append(b["ratio"])
            inside = not inside
        for i in range(len(tirs) - 1, 0, -1):                                # NZ:2063-2064
            m = conv_idxs[i - 1]
            tirs[i - 1][m] = tirs[i - 1][m] & tirs[i]
        # ---- per-segment sample generation (NZ:2067-2146)
        pathes = []
        for k in range(len(converges)):
            start, dk = starts[k], directions[k]
            h_idx = hit_idxs[k]
            n_hit, n_seg = h_idx.numel(), start.shape[0]
            n_pts = 64 if k != 1 else 128
            end = start + dk * 4.5
            if n_hit > 0:
                end = end.index_copy(0, h_idx, intersections[k])
            delta = end - start
            Z = torch.linspace(0, 1, n_pts, device=dev).unsqueeze(0).repeat(n_seg, 1)
            if k == 1 and n_hit > 0:
                # inside the outer mesh: 64 uniform samples to the hit + 2 rounds of 32 SDF-guided ones on the inner field,
                # with the reference's mixed parametrisation (unit parameter z, SDF queried at start + dir * z; NZ:2098-2117)
                with torch.no_grad():
                    s_h = start.detach().index_select(0, h_idx).contiguous()
                    e_h = end.detach().index_select(0, h_idx).contiguous()
                    d_h = dk.detach().index_select(0, h_idx).contiguous()
                    Rh = s_h.shape[0]
                    z = torch.linspace(0, 1, 64, device=dev).unsqueeze(0).expand(Rh, 64).contiguous()
                    p64 = eng.segment_points(s_h, e_h - s_h, z)
                    sdf = eng.sdf_infer(wi.sdf, p64.reshape(-1, 3), wi.planes).reshape(Rh, 64).contiguous()
                    z = eng.upsample_rounds(wi, s_h, d_h, z, sdf, n_new=32, rounds=2)
                    if trace is not None and "z_1" in trace:       # tests: the sample parameters of a recorded trace
                        z = trace["z_1"]
                    Z[h_idx] = z
            if n_hit < n_seg:
                if k == 1:
                    raise RuntimeError("non-zero-thickness trace: a ray of segment 1 has no exit hit after the filter")
                # rays that leave the scene: 64 inverse-depth samples along the direction (NZ:2140-2143)
                miss = infinity_bkgr[k]                                     # [n_seg, 1]: selected in place, no index list
                Z = torch.where(miss, outside_depths(dev).unsqueeze(0), Z)
                delta = torch.where(miss, dk, delta)
            if start.requires_grad or delta.requires_grad:
                pts = start[:, None, :] + delta[:, None, :] * Z[:, :, None]
            else:
                pts = eng.segment_points(start.contiguous(), delta.contiguous(), Z.contiguous())
            pathes.append(pts)
        self._trace_cache = (list(converges), list(conv_idxs))          # see the zero-thickness ray_trace
        return pathes, converges, directions, ior_ratios, infinity_bkgr, gradient_mesh, tirs[0]

    # ------------------------------------------------------------------ NZ:2155-2353
    def _eval_inner(self, pack, inv_s, pts, dirs, dists, params, step, is_train):
        """With apply_occ_loss from occ_loss_step on: the occlusion-probe loss of the inner field (NZ:2222-2230,
        :1580-1608), on the stage-1 machinery (fused SDF query + probe kernel) with the inner operands."""
        if not (self.cfg["apply_occ_loss"] and step is not None and step >= self.cfg["occ_loss_step"]):
            return super()._eval_inner(pack, inv_s, pts, dirs, dists, params, step, is_train)
        a_i, c_i, gerr, occ, pts_in, sdf_in, grad_in, dirs_in, refl_in = _InnerOccFn.apply(pack, inv_s, pts, dirs, dists,
                                                                                          *params)
        self._loss_occ = _InnerOccAdapter(self).compute_occ_loss({"occ_prob": occ, "reflective": refl_in}, pts_in, sdf_in,
                                                                 grad_in, dirs_in, step, prepared=pack[0],
                                                                 perm=getattr(self, "_occ_perm", None))
        return a_i, c_i, gerr

    def render_core(self, rays_o, rays_d, pathes, converges, directions, infinity_bkgr, gradient_mesh, ior_ratios,
                    human_poses=None, cos_anneal_ratio=0.0, step=None, is_train=True, is_nerf=False, prepared=None):
        self._loss_occ = None
        out = super().render_core(rays_o, rays_d, pathes, converges, directions, infinity_bkgr, gradient_mesh, ior_ratios,
                                  human_poses, cos_anneal_ratio=cos_anneal_ratio, step=step, is_train=is_train,
                                  is_nerf=is_nerf, prepared=prepared)
        # zeros(1) before occ_loss_step / when disabled / without inner samples (NZ:1584-1585, :2227-2230)
        out["loss_occ"] = self._loss_occ if self._loss_occ is not None else torch.zeros(1, device=rays_o.device)
        self._loss_occ = None
        return out

    # ------------------------------------------------------------------ NZ:1482-1506
    def render(self, rays_o, rays_d, mask=None, near=None, far=None, human_poses=None, perturb_overwrite=-1,
               cos_anneal_ratio=0.0, is_train=True, step=None, is_nerf=False):
        prepared = self._prepare()
        pathes, converges, directions, ior_ratios, infinity_bkgr, gradient_mesh, tir_mask = \
            self.ray_trace(rays_o, rays_d, mask, prepared=prepared)
        ret = self.render_core(rays_o, rays_d, pathes, converges, directions, infinity_bkgr, gradient_mesh, ior_ratios,
                               human_poses, cos_anneal_ratio=cos_anneal_ratio, step=step, is_train=is_train,
                               is_nerf=is_nerf, prepared=prepared)
        ret["tir_mask"] = tir_mask
        return ret

    @torch.no_grad()
    def nvs(self, pose, K, h, w, chunk=4096):
        from . import feeder
        dev = self.deviation_network_inner.variance.device
        as_t = lambda a: (torch.from_numpy(np.asarray(a, dtype=np.float32)) if not torch.is_tensor(a) else a.float()).to(dev)
        K_, pose_ = as_t(K).unsqueeze(0), as_t(pose).unsqueeze(0)
        batch, rn, _, _ = feeder.construct_ray_batch(torch.zeros(1, 3, h, w, device=dev), K_)
        colors = []
        for ri in range(0, rn, chunk):
            rays_o, rays_d = feeder.world_rays(batch["dirs"][ri:ri + chunk], batch["idxs"][ri:ri + chunk], pose_)
            out = self.render(rays_o.contiguous(), rays_d.contiguous(), None, None, None, None, 0, 0, is_train=False,
                              step=300000)
            colors.append(out["ray_rgb"])
        return torch.cat(colors, 0).reshape(h, w, 3).cpu().numpy()

    def test_step(self, index, step):
        """NZ:1264-1313: as the zero-thickness test_step with the image's foreground mask (src['mask'], optional: ones)
        multiplied into ray_rgb / gt_rgb / loss_rgb (NZ:1297-1299)."""
        if getattr(self, "eval_source", None) is None:
            raise RuntimeError("no eval source attached: call set_eval_source(fn) (the image database is outside the hot path)")
        src = self.eval_source(index)
        rays_o, rays_d = src["rays_o"].float(), F.normalize(src["rays_d"].float(), dim=-1)
        h, w = int(src["h"]), int(src["w"])
        rn, trn = rays_o.shape[0], self.cfg["test_ray_num"]
        outputs = {k: [] for k in self.TEST_KEYS}
        with torch.no_grad():
            for ri in range(0, rn, trn):
                cur = self.render(rays_o[ri:ri + trn].contiguous(), rays_d[ri:ri + trn].contiguous(), None, None, None, None,
                                  0, 0, is_train=False, step=step, is_nerf=self.is_nerf)
                for k in self.TEST_KEYS:
                    outputs[k].append(cur[k].detach())
        outputs = {k: torch.cat(v, 0) for k, v in outputs.items()}
        m = outputs["tir_mask"].float()
        if "mask" in src:
            m = m * torch.as_tensor(src["mask"], device=m.device).float().reshape(-1, 1)
        outputs["loss_rgb"] = self.compute_rgb_loss(outputs["ray_rgb"] * m, src["rgbs"] * m)
        outputs["gt_rgb"] = (src["rgbs"] * m).reshape(h, w, 3)
        outputs["ray_rgb"] = (outputs["ray_rgb"] * m).reshape(h, w, 3)
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
        # the reference's ray tables call it 'mask' (real data, NZ:1073-1081) or 'masks' (synthetic data, NZ:56)
        m = batch.get("mask", batch.get("masks"))
        mask = m.reshape(-1, 1).float() if m is not None else torch.ones_like(rays_d[:, :1])
        out = self.render(batch["rays_o"], rays_d, mask, None, None, None, -1, self.get_anneal_val(step), is_train=True,
                          step=step, is_nerf=self.is_nerf)
        tm = out["tir_mask"].detach() * mask
        out["loss_rgb"] = self.compute_rgb_loss(out["ray_rgb"] * tm, batch["rgbs"] * tm)         # NZ:1364
        return out


name2renderer = {
    "shape": NeROShapeRenderer,
    "stage2": Stage2Renderer,
}

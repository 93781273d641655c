"""Drop-in boundary: the renderer classes of the reference's network/renderer_zerothick.py ("ZT"), backed by the
sm_100a engine (nu_nerf_b200/engine.py).  Same class names, constructor `(cfg, training=True)`, parameter /
state_dict names, `forward(data)`, `render(...)`, `sample_ray(...)`, `render_core(...)` signatures and outputs dict.

What stays in torch: parameter storage, the weight-norm re-parametrisation of the (tiny) weight matrices, the
charbonnier loss on [R,3] and scalar bookkeeping.  Everything per-sample runs in libnunerf_b200.so; there is no
CPU or eager fallback (constructing a renderer without CUDA works -- parameters only -- but rendering raises).
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .field import (SDFNetwork, SingleVarianceNetwork, NeRFNetwork, AppShadingNetwork, InfOutNetwork, WNLinear,
                    PlainLinear)

SPHEREPOT_STAGE1_CFG = {
    # configs/shape/nerf/spherepot.yaml of the reference
    "name": "spherepot", "network": "shape", "database_name": "nerf/spherepot", "apply_occ_loss": True,
    "occ_loss_step": 15000, "is_nerf": True, "zero_thickness": True, "get_mask": False,
    "loss": ["nerf_render", "eikonal", "std", "init_sdf_reg", "occ", "mask", "outer_reg"],
    "val_metric": ["shape_render"], "key_metric_name": "psnr", "eikonal_weight": 0.1, "freeze_inv_s_step": 15000,
    "train_dataset_type": "dummy", "dataset_dir": "./datasets", "optimizer_type": "adam", "lr_type": "warm_up_cos",
    "lr_cfg": {}, "total_step": 200000, "val_interval": 5000, "save_interval": 1000, "train_log_step": 20,
}


def load_default_cfg():
    return dict(SPHEREPOT_STAGE1_CFG)


def load_cfg(path):
    """utils/base_utils.py:319-321."""
    import yaml
    with open(path, "r") as f:
        return yaml.load(f, Loader=yaml.FullLoader)


def linear_to_srgb(x):
    """utils/raw_utils.py:5-12 (used on tiny per-ray tensors only)."""
    eps = torch.finfo(torch.float32).eps
    return torch.where(x <= 0.0031308, 323.0 / 25.0 * x, (211.0 * torch.clamp(x, min=eps) ** (5.0 / 12.0) - 11.0) / 200.0)


def _engine():
    from . import engine
    return engine


class _RenderCoreFn(torch.autograd.Function):
    """ZT:725-793 as one autograd node (see engine.core_forward / core_backward)."""

    @staticmethod
    def forward(ctx, pack, inv_s, *params):
        eng = _engine()
        w, rays_o, rays_d, z_vals, cos_anneal, is_nerf, exp_max, want_inv_s, want_weights = pack
        t, rgb, acc, bkgr, wts = eng.core_forward(w, rays_o, rays_d, z_vals, cos_anneal, is_nerf, exp_max,
                                                  want_weights=want_weights)
        ctx.tape, ctx.w, ctx.n_params, ctx.want_inv_s = t, w, len(params), want_inv_s
        dev = rgb.device
        if t.n_in > 0:
            gerr, trans, met, occ = t.gerr, t.trans[:, None], t.metallic[:, None], t.occ[:, None]
        else:
            gerr, trans, met, occ = (torch.zeros(1, device=dev), torch.zeros(0, 1, device=dev),
                                     torch.zeros(0, 1, device=dev), torch.zeros(0, 1, device=dev))
        spec = t.ls_.head[:, :3].clone()
        if t.n_in > 0:
            aux = (t.pts_in, t.sdf.sdf[:, 0], t.sdf.grad, t.dirs_in, t.refl)
        else:
            aux = tuple(torch.zeros(0, 3, device=dev) if i != 1 else torch.zeros(0, device=dev) for i in range(5))
        ctx.mark_non_differentiable(wts, *aux)
        return (rgb, acc, bkgr, gerr, trans, met, spec, occ, wts) + aux

    @staticmethod
    def backward(ctx, d_rgb, d_acc, d_bkgr, d_gerr, d_trans, d_met, d_spec, d_occ, *_unused):
        eng = _engine()
        t = ctx.tape
        c = lambda x: None if x is None else x.contiguous().float()
        has = t.n_in > 0
        ctx.w.bank.zero_grads()
        g = eng.core_backward(ctx.w, t, c(d_rgb), c(d_acc), c(d_bkgr), c(d_gerr) if has else None,
                              c(d_trans).reshape(-1) if (has and d_trans is not None) else None,
                              c(d_met).reshape(-1) if (has and d_met is not None) else None, c(d_spec), ctx.want_inv_s,
                              d_occ=c(d_occ).reshape(-1) if (has and d_occ is not None) else None)
        # effective-weight gradients -> d weight_v / d weight_g / d bias, added in place to .grad (one launch);
        # autograd therefore receives None for the parameter inputs
        ctx.w.bank.backward()
        ctx.tape = None
        return (None, g.get("inv_s").reshape(()) if ctx.want_inv_s and "inv_s" in g else None,
                *([None] * ctx.n_params))


class _SdfValueFn(torch.autograd.Function):
    """sdf_network.sdf(points) with a backward to the SDF network's parameters (the `sdf_vals` warm-up output of
    ZT:804-807, regularised by InitSDFRegLoss, network/loss.py:115-148, during the first 1000 steps)."""

    @staticmethod
    def forward(ctx, pack, *params):
        eng = _engine()
        w, pts = pack
        M = pts.shape[0]
        xm = eng.P(M, 320, w.planes, pts.device)
        tape = eng.sdf_forward(w.sdf, pts, w.planes, xm)
        ctx.w, ctx.tape, ctx.n = w, tape, len(params)
        return tape.sdf[:, 0].clone()

    @staticmethod
    def backward(ctx, d_sdf):
        eng = _engine()
        w, t = ctx.w, ctx.tape
        M, dev = t.M, t.pts.device
        w.bank.zero_grads()
        dxm = eng.P(M, 320, w.planes, dev, zero=True)                # no gradient arrives through the feature columns
        eng.sdf_backward(w.sdf, t, w.planes, dxm, d_sdf.contiguous().float(), torch.zeros(M, 3, device=dev))
        w.bank.backward()
        ctx.tape = None
        return (None,) + (None,) * ctx.n


class NeROShapeRenderer(nn.Module):
    default_cfg = {
        # standard deviation for opacity density
        "std_net": "default", "std_act": "exp", "inv_s_init": 0.3, "freeze_inv_s_step": None,
        # geometry network
        "sdf_net": "default", "sdf_activation": "none", "sdf_bias": 0.5, "sdf_n_layers": 8, "sdf_freq": 6,
        "sdf_d_out": 257, "geometry_init": True,
        # shader network
        "shader_config": {},
        # sampling strategy
        "n_samples": 64, "n_bg_samples": 32, "inf_far": 1000.0, "n_importance": 64, "up_sample_steps": 4,
        "perturb": 1.0, "anneal_end": 50000, "train_ray_num": 512, "test_ray_num": 1024,
        "clip_sample_variance": True, "is_nerf": False,
        # dataset
        "database_name": "nerf_synthetic/lego/black_800",
        # validation
        "test_downsample_ratio": True, "downsample_ratio": 0.5, "val_geometry": False,
        # losses
        "rgb_loss": "charbonier", "apply_occ_loss": True, "occ_loss_step": 20000, "occ_loss_max_pn": 2048,
        "occ_sdf_thresh": 0.01,
        "fixed_camera": False,
        # B200 engine: "split" (fp32-accurate, 3 MMAs per product) or "bf16" (fast)
        "precision": "split",
    }

    def __init__(self, cfg, training=True):
        super().__init__()
        self.cfg = {**self.default_cfg, **cfg}
        self.is_nerf = self.cfg["is_nerf"]
        if (self.cfg["sdf_n_layers"], self.cfg["sdf_freq"], self.cfg["sdf_d_out"]) != (8, 6, 257):
            raise NotImplementedError("the B200 engine is built for the 8x256 / PE-6 / 257-output SDF network")
        if (self.cfg["n_samples"], self.cfg["n_bg_samples"], self.cfg["n_importance"], self.cfg["up_sample_steps"]) \
                != (64, 32, 64, 4) or not self.cfg["clip_sample_variance"]:
            raise NotImplementedError("sampling kernels are built for 64 + 4x16 + 32 samples with clipped variance")
        self.sdf_network = SDFNetwork(d_out=257, d_in=3, d_hidden=256, n_layers=8, skip_in=(4,), multires=6,
                                      bias=self.cfg["sdf_bias"], scale=1.0, geometric_init=self.cfg["geometry_init"])
        self.deviation_network = SingleVarianceNetwork(init_val=self.cfg["inv_s_init"], activation=self.cfg["std_act"])
        self.outer_nerf = NeRFNetwork(D=8, d_in=4, d_in_view=3, W=256, multires=10, multires_view=4, skips=(4,))
        nn.init.constant_(self.outer_nerf.rgb_linear.bias, np.log(0.5))
        self.color_network = AppShadingNetwork(self.cfg["shader_config"])
        self.infinity_far_bkgr = InfOutNetwork()
        self.sdf_network._query = self._sdf_query
        self.ray_source = None       # callable(step, n) -> dict(rays_o, rays_d, rgbs) ; replaces the image database
        self.eval_source = None      # callable(index) -> rays + ground truth of one test view (see set_eval_source)
        if training:
            self._init_dataset()

    # ------------------------------------------------------------------ data
    def _init_dataset(self):
        """ZT:167-197.  Dataset ingest (image decoding, pose files) is outside the hot path and stays the reference's own
        host code: when this module runs inside the reference tree (the one-line import swap of INTEGRATION.md) and
        cfg['database_name'] is set, the reference's `dataset.database` loads the views and `attach_database` turns
        them into the device-side ray table; anywhere else attach sources with set_ray_source / set_eval_source /
        attach_database."""
        self.ray_source = None
        if not self.cfg.get("database_name"):
            return
        try:
            from dataset.database import parse_database_name, get_database_split     # the reference's package
        except ImportError:
            return
        database = parse_database_name(self.cfg["database_name"], self.cfg.get("dataset_dir"))
        train_ids, test_ids = get_database_split(database)
        self.attach_database(database, train_ids, test_ids)

    def attach_database(self, database, train_ids, test_ids=()):
        """Ray sources from any object with the reference's BaseDatabase interface (get_image -> uint8 [h,w,3], get_pose
        [3,4], get_K [3,3], get_depth -> (depth, mask)): build_imgs_info + _construct_(nerf_)ray_batch + the shuffled
        slicing of train_step (ZT:20-55, :199-255, :447-466) on the device-side feeder, and the per-view source of
        test_step (ZT:397-411, without the optional test_downsample_ratio blur).  The tables are built on first use, on
        the device the module lives on then."""
        self.database = database
        self.train_ids, self.test_ids = np.asarray(train_ids), list(test_ids)
        self.train_num, self.test_num = len(self.train_ids), len(self.test_ids)
        self._tables = None

        def info(ids, with_depth):
            imgs = np.stack([database.get_image(i) for i in ids], 0).astype(np.float32) / 255.0     # color_map_forward
            out = {"imgs": torch.from_numpy(imgs).permute(0, 3, 1, 2),
                   "Ks": torch.from_numpy(np.stack([database.get_K(i) for i in ids], 0).astype(np.float32)),
                   "poses": torch.from_numpy(np.stack([database.get_pose(i) for i in ids], 0).astype(np.float32))}
            if self.is_nerf or with_depth:
                dm = [database.get_depth(i) for i in ids]
                out["depths"] = torch.from_numpy(np.stack([d_[0] for d_ in dm], 0).astype(np.float32))
                out["masks"] = torch.from_numpy(np.stack([d_[1] for d_ in dm], 0))
            return out

        def tables():
            if self._tables is None:
                from . import feeder
                dev = self.deviation_network.variance.device
                tr = {k: v.to(dev) for k, v in info(self.train_ids, False).items()}
                if self.is_nerf:
                    batch, _, _, _ = feeder.construct_nerf_ray_batch(tr["imgs"], tr["Ks"], tr["poses"], tr["masks"])
                    train = feeder.DeviceRayFeeder(batch, perm_device="cpu")
                else:
                    batch, _, _, _ = feeder.construct_ray_batch(tr["imgs"], tr["Ks"])
                    train = feeder.DeviceRayFeeder(batch, poses=tr["poses"], perm_device="cpu")
                ev = None
                if self.test_num:
                    te = {k: v.to(dev) for k, v in info(self.test_ids, True).items()}
                    ev = feeder.image_eval_source(te["imgs"], te["Ks"], te["poses"], is_nerf=self.is_nerf,
                                                  depths=te["depths"], masks=te["masks"])
                self._tables = (train, ev)
            return self._tables
        self.ray_source = lambda step, n: tables()[0](step, n)
        if self.test_num:
            self.eval_source = lambda index: tables()[1](index)

    def set_ray_source(self, fn):
        self.ray_source = fn

    # ------------------------------------------------------------------ helpers identical to the reference
    def set_eval_source(self, fn):
        """fn(index) -> {'rays_o' [h*w,3], 'rays_d' [h*w,3], 'rgbs' [h*w,3], 'h', 'w'[, 'gt_depth' [h,w], 'gt_mask' [h,w]]}:
        the rays and ground truth of test view `index`, already at the evaluation resolution (what ZT:398-411 reads from the
        image database and down-samples; `nu_nerf_b200.feeder.image_eval_source` builds it from image tensors)."""
        self.eval_source = fn

    def get_anneal_val(self, step):
        if self.cfg["anneal_end"] < 0:
            return 1.0
        return np.min([1.0, step / self.cfg["anneal_end"]])

    @staticmethod
    def near_far_from_sphere(rays_o, rays_d):
        a = torch.sum(rays_d ** 2, dim=-1, keepdim=True)
        b = 2.0 * torch.sum(rays_o * rays_d, dim=-1, keepdim=True)
        mid = 0.5 * (-b) / a
        return torch.clamp(mid - 1.0, min=1e-3), mid + 1.0

    def compute_rgb_loss(self, rgb_pr, rgb_gt):
        kind = self.cfg["rgb_loss"]
        if kind == "l2":
            return torch.sum((rgb_pr - rgb_gt) ** 2, -1)
        if kind == "l1":
            return torch.sum(F.l1_loss(rgb_pr, rgb_gt, reduction="none"), -1)
        if kind == "smooth_l1":
            return torch.sum(F.smooth_l1_loss(rgb_pr, rgb_gt, reduction="none", beta=0.25), -1)
        if kind == "charbonier":
            return torch.sqrt(torch.sum((rgb_gt - rgb_pr) ** 2, dim=-1) + 0.001)
        raise NotImplementedError

    # ------------------------------------------------------------------ weights
    def effective_weights(self):
        """{reference state_dict-style name -> effective fp32 tensor}: weight-norm applied (`W = g v / |v|`)."""
        Wd = {}
        for name, mod in self.named_modules():
            if isinstance(mod, WNLinear):
                Wd[name + ".weight"] = mod.effective_weight()
                Wd[name + ".bias"] = mod.bias
            elif isinstance(mod, PlainLinear):
                Wd[name + ".weight"] = mod.weight
                Wd[name + ".bias"] = mod.bias
        Wd["deviation_network.variance"] = self.deviation_network.variance
        Wd["color_network.FG_LUT"] = self.color_network.FG_LUT
        return Wd

    def _planes(self):
        p = self.cfg["precision"]
        if p not in ("split", "bf16"):
            raise NotImplementedError(f"precision {p}")
        return 2 if p == "split" else 1

    def _prepare(self):
        """Operands of this step: persistent WeightBank (rebuilt only if the parameter storage or the precision
        changed) refreshed from the current parameter values by one kernel launch."""
        dev = self.deviation_network.variance.device
        if not torch.cuda.is_available() or dev.type != "cuda":
            raise RuntimeError("nu_nerf_b200 renders on a CUDA device only (no CPU fallback): move the module with .cuda()")
        from ._lib import require_current_device
        require_current_device(self.deviation_network.variance)       # the C entry points launch on the current device's stream
        eng = _engine()
        key = (self._planes(),) + tuple(p.data_ptr() for p in self.parameters())
        if getattr(self, "_w", None) is None or self._w_key != key:
            self._w, self._w_key = eng.Stage1Weights(self, self._planes(), dev), key
        self._w.refresh()
        return self._w

    @torch.no_grad()
    def _sdf_query(self, x):
        eng = _engine()
        w = self._prepare()
        shape = x.shape[:-1]
        pts = x.reshape(-1, 3).float().contiguous()
        return eng.sdf_infer(w.sdf, pts, w.planes).reshape(*shape, 1)

    # ------------------------------------------------------------------ ZT:572-612
    def sample_ray(self, rays_o, rays_d, near, far, perturb, uniforms=None, prepared=None, trace=None, sphere=False):
        eng = _engine()
        w = prepared if prepared is not None else self._prepare()
        R, dev = rays_o.shape[0], rays_o.device
        U0 = U1 = None
        if perturb > 0:
            if uniforms is not None:
                U0, U1 = (u.contiguous().float() for u in uniforms)
            else:
                U0 = torch.rand([R, 1], device=dev)       # same draw order as ZT:585, :591
                U1 = torch.rand([R, 32], device=dev)
        return eng.sample_ray(w.sdf, w.inv_s, w.planes, rays_o, rays_d, near, far, perturb > 0, U0, U1, sphere=sphere,
                              trace=trace)

    @torch.no_grad()
    def count_inner(self, rays_o, rays_d, z_vals):
        """Number of samples whose mid-point lies inside the unit sphere (the inner mask of render_core, ZT:730-741) as a
        0-d device tensor, without rendering: the length `gradient_error` WILL have.  A chunked / ray-sharded trainer
        needs it before any chunk is differentiated (nu_nerf_b200/dist.py)."""
        eng = _engine()
        R, S = z_vals.shape
        counts = torch.empty(R, dtype=torch.int32, device=z_vals.device)
        eng.call("nunerf_inner_counts", rays_o.contiguous().float().data_ptr(), rays_d.contiguous().float().data_ptr(),
                 z_vals.contiguous().float().data_ptr(), R, S, counts.data_ptr())
        return counts.sum()

    # ------------------------------------------------------------------ ZT:725-820
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
        (rgb, acc, bkgr, gerr, trans, met, spec, occ, weights, pts_in, sdf_in, grad_in, dirs_in,
         refl_in) = _RenderCoreFn.apply(pack, inv_s, *params)
        inv_s_c = inv_s.clip(1e-6, 1e6)
        if frozen:
            inv_s_c = inv_s_c.detach()
        has_inner = trans.shape[0] > 0
        outputs = {
            "ray_rgb": rgb, "gradient_error": gerr, "acc": acc, "color_bkgr": bkgr,
            "color_spec": linear_to_srgb(torch.exp(torch.clamp(spec, max=exp_max))),
            "std": torch.mean(1.0 / inv_s_c) if has_inner else torch.zeros(1, device=rgb.device),
        }
        if has_inner:
            outputs["transmission"] = trans
            outputs["metallic"] = met
        if self.cfg["apply_occ_loss"]:
            if has_inner and step is not None:
                outputs["loss_occ"] = self.compute_occ_loss({"occ_prob": occ, "reflective": refl_in}, pts_in, sdf_in,
                                                            grad_in, dirs_in, step, prepared=w, perm=occ_perm)
            else:
                outputs["loss_occ"] = torch.zeros(1, device=rgb.device)
        if step is not None and step < 1000:
            # warm-up outputs of ZT:804-807: every sample inside radius 1.2 and its (differentiable) SDF value.  Only the
            # first 1000 of 200 000 steps take this branch; the mid points are rebuilt with torch indexing (ZT:730-736)
            zf = z_vals.float()
            dist = torch.cat([zf[:, 1:] - zf[:, :-1], zf[:, -1:] - zf[:, -2:-1]], -1)
            pts_all = rays_o.float()[:, None, :] + rays_d.float()[:, None, :] * (zf + dist * 0.5)[..., None]
            mask = torch.norm(pts_all, dim=-1) < 1.2
            sdf_pts = pts_all[mask].contiguous()
            outputs["sdf_pts"] = sdf_pts
            outputs["sdf_vals"] = _SdfValueFn.apply((w, sdf_pts), *params) if sdf_pts.shape[0] > 0 else \
                torch.zeros(0, device=rgb.device)
        if not is_train:
            outputs.update(self.compute_validation_info(z_vals, rays_o, rays_d, weights, human_poses, step, prepared=w))
        return outputs

    # ------------------------------------------------------------------ ZT:695-723, field.py:501-554
    def compute_occ_loss(self, occ_info, points, sdf, gradients, dirs, step, prepared=None, perm=None, sync_free=True):
        """L1 between the predicted occlusion probability and the hit probability of a 64 + 16 sample SDF probe along
        the reflected ray, on (at most occ_loss_max_pn) surface samples.  torch glue on <= 2048-row tensors around the
        fused SDF-inference kernel; `perm` injects the torch.randperm draw of ZT:710 (parity tests).

        Default (perm is None and sync_free): the subset is drawn WITHOUT a host round trip -- every candidate gets a
        uniform random key, the occ_loss_max_pn smallest keys are kept (a uniform random subset without replacement,
        like randperm[:max_pn]; all candidates when there are fewer) and padding slots are masked out of the mean.  The
        reference's `int(mask.sum())` / `nonzero` are host round trips in the middle of every step from occ_loss_step on
        (92 % of training); they are avoided so that the launch queue never drains there."""
        dev = points.device
        if step < self.cfg["occ_loss_step"]:
            return torch.zeros(1, device=dev)
        occ_prob, reflective = occ_info["occ_prob"], occ_info["reflective"]
        mask = (torch.norm(points, dim=-1) < 0.999) & (torch.sum(gradients * dirs, -1) < 0) & \
               (torch.abs(sdf) < self.cfg["occ_sdf_thresh"])
        max_pn = self.cfg["occ_loss_max_pn"]
        if perm is None and sync_free:
            k = min(max_pn, points.shape[0])
            keys = torch.where(mask, torch.rand(points.shape[0], device=dev), torch.full((), 2.0, device=dev))
            vals, idx = torch.topk(keys, k, largest=False, sorted=False)
            valid = vals < 1.5
            # padding slots probe a harmless ray from the origin (their weight in the mean is zero)
            pts_sel = torch.where(valid[:, None], points[idx], torch.zeros(1, 3, device=dev))
            e_x = torch.eye(1, 3, device=dev)                    # [[1, 0, 0]] built on the device: a host->device copy
                                                                 # (torch.tensor / item assignment) would drain the stream
            dirs_sel = torch.where(valid[:, None], reflective[idx].detach(), e_x).contiguous()
            w = prepared if prepared is not None else self._prepare()
            occ_gt = self.occ_probability(pts_sel.contiguous(), dirs_sel, w)
            diff = (occ_prob[idx] - occ_gt).abs() * valid[:, None]
            return diff.sum() / torch.clamp(valid.sum() * occ_prob.shape[-1], min=1)
        n = int(mask.sum())
        if n > max_pn:
            indices = torch.nonzero(mask)[:, 0]
            idx = perm.to(dev) if perm is not None else torch.randperm(indices.shape[0], device=dev)
            indices = indices[idx[:max_pn]]
            mask = torch.zeros_like(mask)
            mask[indices] = True
        if n == 0:
            return torch.zeros(1, device=dev)
        w = prepared if prepared is not None else self._prepare()
        occ_gt = self.occ_probability(points[mask], reflective[mask].detach(), w)
        return F.l1_loss(occ_prob[mask], occ_gt)

    @torch.no_grad()
    def occ_probability(self, pts, dirs, w, sn0=64, sn1=16, use_kernels=True):
        """get_intersection (field.py:524-554): probability that the ray pts + t dirs hits the surface before it
        leaves the unit sphere.  pts must lie inside radius 0.999 (the caller's mask)."""
        eng = _engine()
        inv_s = torch.exp(self.deviation_network.variance.detach() * 10.0)
        dtx = torch.sum(pts * dirs, dim=-1, keepdim=True)
        xtx = torch.sum(pts ** 2, dim=-1, keepdim=True)
        max_dist = -dtx + torch.sqrt(dtx ** 2 - xtx + 1 + 1e-6)                 # get_sphere_intersection :458-464

        P = pts.shape[0]
        dev = pts.device
        if use_kernels and sn0 <= 64 and sn1 <= 32:
            # kernels: points, fused SDF query, one warp-per-ray probe kernel per pass (csrc/sampling.cu)
            pts_c, dirs_c = pts.contiguous().float(), dirs.contiguous().float()
            inv_s_dev = inv_s.reshape(1).float().contiguous()
            u_tab = eng._tables(dev)[1].get(sn1)
            if u_tab is None:
                u_tab = torch.linspace(0.5 / sn1, 1.0 - 0.5 / sn1, sn1, device=dev)
            z = (max_dist * torch.linspace(0, 1, sn0, device=dev).unsqueeze(0)).contiguous()
            sdf = eng.sdf_infer(w.sdf, eng.segment_points(pts_c, dirs_c, z).reshape(-1, 3), w.planes).reshape(P, sn0)
            z_new = torch.empty(P, sn1, device=dev)
            eng.call("nunerf_probe_weights", z.data_ptr(), sdf.contiguous().data_ptr(), P, sn0, inv_s_dev.data_ptr(), sn1,
                     u_tab.data_ptr(), z_new.data_ptr(), None)
            sdf2 = eng.sdf_infer(w.sdf, eng.segment_points(pts_c, dirs_c, z_new).reshape(-1, 3), w.planes).reshape(P, sn1)
            wsum = torch.empty(P, device=dev)
            eng.call("nunerf_probe_weights", z_new.data_ptr(), sdf2.contiguous().data_ptr(), P, sn1, inv_s_dev.data_ptr(), 0,
                     None, None, wsum.data_ptr())
            return wsum[:, None]

        def weights(z):                                                        # get_weights :501-521
            p = (z.unsqueeze(-1) * dirs.unsqueeze(-2) + pts.unsqueeze(-2)).reshape(-1, 3).contiguous()
            sdf = eng.sdf_infer(w.sdf, p, w.planes).reshape(z.shape)
            ps, ns, pz, nz = sdf[:, :-1], sdf[:, 1:], z[:, :-1], z[:, 1:]
            mid = (ps + ns) * 0.5
            cos = (ns - ps) / (nz - pz + 1e-5)
            surf = cos < 0
            cos = torch.clamp(cos, max=0)
            dist = nz - pz
            pc = torch.sigmoid((mid - cos * dist * 0.5) * inv_s)
            nc = torch.sigmoid((mid + cos * dist * 0.5) * inv_s)
            alpha = (pc - nc + 1e-5) / (pc + 1e-5) * surf.float()
            T = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1.0 - alpha + 1e-7], -1), -1)[:, :-1]
            return alpha * T

        # wider probes (the eval path uses 128 + 9 samples): torch glue around the fused SDF query
        z = max_dist * torch.linspace(0, 1, sn0, device=pts.device).unsqueeze(0)
        wts = weights(z) + 1e-5                                                # sample_pdf(det) field.py:468-498
        pdf = wts / torch.sum(wts, -1, keepdim=True)
        cdf = torch.cat([torch.zeros_like(pdf[:, :1]), torch.cumsum(pdf, -1)], -1)
        u = torch.linspace(0.5 / sn1, 1.0 - 0.5 / sn1, sn1, device=pts.device).expand(cdf.shape[0], sn1).contiguous()
        inds = torch.searchsorted(cdf, u, right=True)
        lo, hi = torch.clamp(inds - 1, min=0), torch.clamp(inds, max=cdf.shape[-1] - 1)
        c0, c1 = torch.gather(cdf, 1, lo), torch.gather(cdf, 1, hi)
        b0, b1 = torch.gather(z, 1, lo), torch.gather(z, 1, hi)
        den = c1 - c0
        den = torch.where(den < 1e-5, torch.ones_like(den), den)
        z_new = b0 + (u - c0) / den * (b1 - b0)
        return torch.sum(weights(z_new), -1, keepdim=True)

    def compute_validation_info(self, z_vals, rays_o, rays_d, weights, human_poses, step, prepared=None):
        """ZT:636-655: depth = sum w z ; normal = (normalize(grad sdf(o + depth d)) + 1) / 2 inside the unit sphere; the
        shading network's intermediate buffers at the depth point (field.py:749-772) and the probed occlusion
        probability `occ_prob_gt` (get_intersection with 128 + 9 samples, field.py:524-554), all masked to the sphere."""
        eng = _engine()
        w = prepared if prepared is not None else self._prepare()
        with torch.no_grad():
            depth = torch.sum(weights * z_vals, -1, keepdim=True)
            points = (depth * rays_d + rays_o).contiguous()
            R, dev = points.shape[0], points.device
            t = eng.inner_tape(points, rays_d.contiguous().float(), None, 0.0, self.color_network.cfg["light_exp_max"])
            t.xm = eng.P(R, 320, w.planes, dev)
            eng.f32_to_planes(points, t.xm, R, 3, 64, col=256)
            t.sdf = eng.sdf_forward(w.sdf, points, w.planes, t.xm)
            inner = torch.norm(points, dim=-1, keepdim=True) <= 1.0
            outputs = {"depth": depth, "normal": ((F.normalize(t.sdf.grad, dim=-1) + 1.0) * 0.5) * inner}
            eng.shade_forward(w, t, t.sdf.grad)
            occ_gt = torch.zeros(R, 1, device=dev)
            inside = (torch.norm(points, dim=-1) < 0.999).nonzero().squeeze(1)      # one host sync for the mask
            if inside.numel() > 0:
                occ_gt[inside] = self.occ_probability(points[inside], t.refl[inside].contiguous(), w, sn0=128, sn1=9)
            outputs["occ_prob_gt"] = occ_gt
            for k, v in eng.shading_buffers(w, t, t.exp_max).items():
                outputs[k] = v * inner
        return outputs

    # ------------------------------------------------------------------ ZT:614-634
    def render(self, rays_o, rays_d, near, far, human_poses=None, perturb_overwrite=-1, cos_anneal_ratio=0.0,
               is_train=True, step=None, is_nerf=False, uniforms=None, occ_perm=None):
        perturb = self.cfg["perturb"]
        if perturb_overwrite >= 0:
            perturb = perturb_overwrite
        prepared = self._prepare()
        z_vals = self.sample_ray(rays_o, rays_d, near, far, perturb, uniforms=uniforms, prepared=prepared)
        return self.render_core(rays_o, rays_d, z_vals, human_poses, cos_anneal_ratio=cos_anneal_ratio, step=step,
                                is_train=is_train, is_nerf=is_nerf, prepared=prepared, occ_perm=occ_perm)

    # ------------------------------------------------------------------ ZT:278-311
    @torch.no_grad()
    def nvs(self, pose, K, h, w, chunk=8192):
        """Novel-view synthesis: `pose` [3,4] world-to-camera, `K` [3,3] (numpy or tensors) -> ray_rgb [h,w,3] numpy, with
        the reference's eval settings (sphere-bounded near/far, no perturbation, cos_anneal 0, step 300000)."""
        from . import feeder
        dev = self.deviation_network.variance.device
        as_t = lambda a: (torch.from_numpy(np.asarray(a, dtype=np.float32)) if not torch.is_tensor(a) else a.float()).to(dev)
        K_, pose_ = as_t(K).unsqueeze(0), as_t(pose).unsqueeze(0)
        batch, rn, _, _ = feeder.construct_ray_batch(torch.zeros(1, 3, h, w, device=dev), K_)
        colors = []
        for ri in range(0, rn, chunk):
            rays_o, rays_d = feeder.world_rays(batch["dirs"][ri:ri + chunk], batch["idxs"][ri:ri + chunk], pose_)
            near, far = self.near_far_from_sphere(rays_o, rays_d)
            out = self.render(rays_o.contiguous(), rays_d.contiguous(), near, far, None, 0, 0, is_train=False, step=300000)
            colors.append(out["ray_rgb"])
        return torch.cat(colors, 0).reshape(h, w, 3).cpu().numpy()

    # ------------------------------------------------------------------ ZT:846-864, field.py:779-783
    @torch.no_grad()
    def predict_materials(self, xyz=None, batch_size=8192 * 8):
        """metallic / roughness / albedo of the material predictors at surface points.  The reference reads the vertices
        of `data/meshes/{name}-300000.ply` (ZT:847-849); here `xyz` is an [N,3] tensor / array of points, or a mesh file
        path (.ply / .npz, nu_nerf_b200.tracer.load_mesh), or None for the reference's default path."""
        eng = _engine()
        w = self._prepare()
        dev = self.deviation_network.variance.device
        if xyz is None or isinstance(xyz, str):
            from .tracer import load_mesh
            xyz = load_mesh(xyz if xyz is not None else f"data/meshes/{self.cfg['name']}-300000.ply")[0]
        pts_all = torch.as_tensor(np.asarray(xyz, dtype=np.float32) if not torch.is_tensor(xyz) else xyz).float().to(dev)
        res = {"metallic": [], "roughness": [], "albedo": []}
        for vi in range(0, pts_all.shape[0], batch_size):
            pts = pts_all[vi:vi + batch_size].contiguous()
            M = pts.shape[0]
            xm = eng.P(M, 320, w.planes, dev)
            eng.f32_to_planes(pts, xm, M, 3, 64, col=256)
            eng.sdf_forward(w.sdf, pts, w.planes, xm)                       # feature vector -> xm[:, :256]
            for key, name, n in (("metallic", "metallic_predictor", 1), ("roughness", "roughness_predictor", 1),
                                 ("albedo", "albedo_predictor", 3)):
                head = eng.pred_forward(w.pred[name], xm, M, 320, w.planes).head[:, :n]
                res[key].append(torch.sigmoid(head).cpu().numpy())
        return {k: np.concatenate(v, 0) for k, v in res.items()}

    # ------------------------------------------------------------------ ZT:447-466
    def train_step(self, step):
        if self.ray_source is None:
            raise RuntimeError("no ray source attached: call set_ray_source(fn) (dataset ingest is outside the hot path)")
        rn = self.cfg["train_ray_num"]
        batch = self.ray_source(step, rn)
        self._last_batch = batch
        rays_o = batch["rays_o"]
        rays_d = F.normalize(batch["rays_d"], dim=-1)
        if self.is_nerf:
            near = torch.full((rays_o.shape[0], 1), 0.8, device=rays_o.device)
            far = torch.full((rays_o.shape[0], 1), 4.5, device=rays_o.device)
        else:
            near, far = self.near_far_from_sphere(rays_o, rays_d)
        outputs = self.render(rays_o, rays_d, near, far, None, -1, self.get_anneal_val(step), is_train=True, step=step,
                              is_nerf=self.is_nerf, uniforms=batch.get("uniforms"))
        outputs["loss_rgb"] = self.compute_rgb_loss(outputs["ray_rgb"], batch["rgbs"])
        return outputs

    # ------------------------------------------------------------------ ZT:397-445
    TEST_KEYS = ("ray_rgb", "gradient_error", "normal", "depth", "diffuse_albedo", "diffuse_light", "diffuse_color",
                 "refraction_light", "specular_albedo", "specular_light", "specular_color", "specular_ref",
                 "transmission_weight", "roughness", "occ_prob", "indirect_light", "occ_prob_gt")

    def test_step(self, index, step):
        """One test view in chunks of cfg['test_ray_num'] rays through render(..., perturb 0, anneal 0, is_train=False);
        same output keys and shapes as the reference (ray_rgb / gt_rgb [h,w,3], the buffers [h*w, .], loss_rgb [h*w])."""
        if self.eval_source is None:
            raise RuntimeError("no eval source attached: call set_eval_source(fn) (the image database is outside the hot path)")
        src = self.eval_source(index)
        rays_o, rays_d = src["rays_o"].float(), F.normalize(src["rays_d"].float(), dim=-1)
        h, w = int(src["h"]), int(src["w"])
        rn, trn = rays_o.shape[0], self.cfg["test_ray_num"]
        outputs = {k: [] for k in self.TEST_KEYS}
        with torch.no_grad():
            for ri in range(0, rn, trn):
                o, d = rays_o[ri:ri + trn].contiguous(), rays_d[ri:ri + trn].contiguous()
                if self.is_nerf:                                            # _process_nerf_ray_batch ZT:363-374
                    near = torch.full((o.shape[0], 1), 0.8, device=o.device)
                    far = torch.full((o.shape[0], 1), 4.5, device=o.device)
                else:
                    near, far = self.near_far_from_sphere(o, d)
                cur = self.render(o, d, near, far, None, 0, 0, is_train=False, step=step, is_nerf=self.is_nerf)
                for k in self.TEST_KEYS:
                    outputs[k].append(cur[k].detach())
        outputs = {k: torch.cat(v, 0) for k, v in outputs.items()}
        outputs["loss_rgb"] = self.compute_rgb_loss(outputs["ray_rgb"], src["rgbs"])
        outputs["gt_rgb"] = src["rgbs"].reshape(h, w, 3)
        outputs["ray_rgb"] = outputs["ray_rgb"].reshape(h, w, 3)
        for k in ("gt_depth", "gt_mask"):                                   # used by the evaluation metrics (metrics.py:55-84)
            if k in src:
                outputs[k] = torch.as_tensor(src[k]).unsqueeze(-1)
        self.zero_grad()
        return outputs

    def forward(self, data):
        step = data["step"]
        if "eval" in data:
            return self.test_step(data["index"], step)
        return self.train_step(step)


from .renderer_stage2 import Stage2Renderer  # noqa: E402  (ZT:868-2011)

name2renderer = {
    "shape": NeROShapeRenderer,
    "stage2": Stage2Renderer,
}

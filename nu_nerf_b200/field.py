"""Parameter containers of the NU-NeRF fields, mirroring the reference's module tree.

Only *parameters* live here: names, shapes, registration order and initialisation follow
/root/reference/network/field.py so that (a) reference checkpoints load with load_state_dict and
(b) `torch.manual_seed(s); NeROShapeRenderer(cfg)` yields bit-identical initial weights to the
reference (the constructors consume the torch RNG in the same order).  All arithmetic on these
parameters is done by the CUDA engine (nu_nerf_b200/engine.py); there is no torch forward here.

Reference: SDFNetwork field.py:64-131, SingleVarianceNetwork :191-195, NeRFNetwork :212-263,
make_predictor :371-408, AppShadingNetwork.__init__ :569-616, InfOutNetwork :1020-1036,
IoRNetwork :1046-1059.
"""

import numpy as np
import torch
import torch.nn as nn

from .fg_lut import load_fg_lut


class WNLinear(nn.Module):
    """A weight-normalised dense layer stored the way nn.utils.weight_norm(dim=0) stores it:
    parameters `bias`, `weight_g` [out,1], `weight_v` [out,in] (registered in that order)."""

    def __init__(self, lin: nn.Linear):
        super().__init__()
        self.in_features, self.out_features = lin.in_features, lin.out_features
        w = lin.weight.detach()
        self.bias = nn.Parameter(lin.bias.detach().clone())
        self.weight_g = nn.Parameter(w.norm(dim=1, keepdim=True).clone())
        self.weight_v = nn.Parameter(w.clone())

    def effective_weight(self):
        return self.weight_g * self.weight_v / self.weight_v.norm(dim=1, keepdim=True)


class PlainLinear(nn.Module):
    """nn.Linear storage (weight, bias) without a forward."""

    def __init__(self, in_features, out_features):
        super().__init__()
        lin = nn.Linear(in_features, out_features)
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(lin.weight.detach().clone())
        self.bias = nn.Parameter(lin.bias.detach().clone())


class _Slot(nn.Module):
    """Parameter-free placeholder keeping nn.Sequential indices aligned with the reference."""


def pe_dim(multires, d=3):
    return d * (1 + 2 * multires)


class SDFNetwork(nn.Module):
    """field.py:64-131 (d_in 3, multires 6, 8 hidden x 256, skip_in (4,), geometric init, weight norm)."""

    def __init__(self, d_in=3, d_out=257, d_hidden=256, n_layers=8, skip_in=(4,), multires=6, bias=0.5,
                 scale=1.0, geometric_init=True, inside_outside=False):
        super().__init__()
        dims = [d_in] + [d_hidden] * n_layers + [d_out]
        if multires > 0:
            dims[0] = pe_dim(multires, d_in)
        self.dims, self.skip_in, self.multires, self.scale = dims, tuple(skip_in), multires, scale
        self.num_layers = len(dims)
        for l in range(self.num_layers - 1):
            out_dim = dims[l + 1] - dims[0] if (l + 1) in self.skip_in else dims[l + 1]
            lin = nn.Linear(dims[l], out_dim)
            if geometric_init:
                if l == self.num_layers - 2:
                    sign = -1.0 if inside_outside else 1.0
                    nn.init.normal_(lin.weight, mean=sign * np.sqrt(np.pi) / np.sqrt(dims[l]), std=0.0001)
                    nn.init.constant_(lin.bias, -sign * bias)
                elif multires > 0 and l == 0:
                    nn.init.constant_(lin.bias, 0.0)
                    nn.init.constant_(lin.weight[:, 3:], 0.0)
                    nn.init.normal_(lin.weight[:, :3], 0.0, np.sqrt(2) / np.sqrt(out_dim))
                elif multires > 0 and l in self.skip_in:
                    nn.init.constant_(lin.bias, 0.0)
                    nn.init.normal_(lin.weight, 0.0, np.sqrt(2) / np.sqrt(out_dim))
                    nn.init.constant_(lin.weight[:, -(dims[0] - 3):], 0.0)
                else:
                    nn.init.constant_(lin.bias, 0.0)
                    nn.init.normal_(lin.weight, 0.0, np.sqrt(2) / np.sqrt(out_dim))
            setattr(self, "lin" + str(l), WNLinear(lin))

    def layers(self):
        return [getattr(self, "lin" + str(l)) for l in range(self.num_layers - 1)]

    # the renderer that owns this network installs the engine query (prepared weights live there)
    _query = None

    def sdf(self, x):
        """SDFNetwork.sdf (field.py:152): [N,3] -> [N,1], no-grad engine query (extract_mesh_stage1.py:36)."""
        if self._query is None:
            raise RuntimeError("SDFNetwork.sdf needs the owning renderer's engine (construct it through a renderer)")
        return self._query(x)


class SingleVarianceNetwork(nn.Module):
    """field.py:191-208 -- inv_s = exp(10 * variance)."""

    def __init__(self, init_val, activation="exp"):
        super().__init__()
        if activation != "exp":
            raise NotImplementedError
        self.act = activation
        self.register_parameter("variance", nn.Parameter(torch.tensor(init_val)))


class NeRFNetwork(nn.Module):
    """field.py:212-263 (NeRF++ background field: D 8, W 256, PE-10 on 4-D points, PE-4 on views)."""

    def __init__(self, D=8, W=256, d_in=4, d_in_view=3, multires=10, multires_view=4, skips=(4,)):
        super().__init__()
        self.D, self.W, self.skips = D, W, tuple(skips)
        self.input_ch = pe_dim(multires, d_in)
        self.input_ch_view = pe_dim(multires_view, d_in_view)
        self.pts_linears = nn.ModuleList(
            [PlainLinear(self.input_ch, W)] +
            [PlainLinear(W, W) if i not in self.skips else PlainLinear(W + self.input_ch, W) for i in range(D - 1)])
        self.views_linears = nn.ModuleList([PlainLinear(self.input_ch_view + W, W // 2)])
        self.feature_linear = PlainLinear(W, W)
        self.alpha_linear = PlainLinear(W, 1)
        self.rgb_linear = PlainLinear(W // 2, 3)


def make_predictor(feats_dim, output_dim, activation="sigmoid", exp_max=0.0):
    """field.py:371-408 -- four weight-normed layers at Sequential indices 0,2,4,6."""
    if activation not in ("sigmoid", "exp", "none", "relu"):
        raise NotImplementedError
    run_dim = 256
    mods = []
    for i, (a, b) in enumerate([(feats_dim, run_dim), (run_dim, run_dim), (run_dim, run_dim), (run_dim, output_dim)]):
        mods += [WNLinear(nn.Linear(a, b)), _Slot()]
    seq = nn.Sequential(*mods)
    seq.activation, seq.exp_max = activation, exp_max
    return seq


class AppShadingNetwork(nn.Module):
    """field.py:557-616 (parameters only)."""
    refrac_exp_max = None          # the refraction-light head shares light_exp_max (field.py:602)
    default_cfg = {
        "human_light": False, "sphere_direction": False, "light_pos_freq": 6, "inner_init": -0.95,
        "roughness_init": 0.0, "metallic_init": 0.0, "light_exp_max": 3.0, "refrac_freq": 6,
    }

    def __init__(self, cfg):
        super().__init__()
        self.cfg = {**self.default_cfg, **cfg}
        if self.cfg["human_light"]:
            raise NotImplementedError("the human_light shading variant is not on the B200 path")
        feats_dim = 256
        self.metallic_predictor = make_predictor(feats_dim + 3, 1)
        if self.cfg["metallic_init"] != 0:
            nn.init.constant_(self.metallic_predictor[-2].bias, self.cfg["metallic_init"])
        self.roughness_predictor = make_predictor(feats_dim + 3, 1)
        if self.cfg["roughness_init"] != 0:
            nn.init.constant_(self.roughness_predictor[-2].bias, self.cfg["roughness_init"])
        self.albedo_predictor = make_predictor(feats_dim + 3, 3)
        self.register_buffer("FG_LUT", torch.from_numpy(load_fg_lut()))
        pos_dim = pe_dim(self.cfg["light_pos_freq"])
        dir_dim = pe_dim(6)
        rf = pe_dim(self.cfg["refrac_freq"])
        exp_max = self.cfg["light_exp_max"]
        # sphere_direction (field.py:594-597): the direct light also sees IDE(exit point of the ray on the unit sphere)
        self.outer_light = make_predictor(72 * 2 if self.cfg["sphere_direction"] else 72, 3, activation="exp",
                                          exp_max=exp_max)
        nn.init.constant_(self.outer_light[-2].bias, np.log(0.5))
        self.inner_light = make_predictor(pos_dim + 72, 3, activation="exp", exp_max=exp_max)
        nn.init.constant_(self.inner_light[-2].bias, np.log(0.5))
        self.inner_weight = make_predictor(pos_dim + dir_dim, 1, activation="none")
        nn.init.constant_(self.inner_weight[-2].bias, self.cfg["inner_init"])
        self.transmisstion_weight = make_predictor(feats_dim + 3, 1)
        self.iors = make_predictor(feats_dim + 3, 1)
        self.refrac_light = make_predictor(rf + rf, 3, activation="exp",
                                           exp_max=exp_max if self.refrac_exp_max is None else self.refrac_exp_max)
        nn.init.constant_(self.refrac_light[-2].bias, np.log(0.5))


class AppShadingNetwork_SpecInner(AppShadingNetwork):
    """field.py:1320-1379 -- the inner-field shader of the non-zero-thickness stage 2 (network/renderer.py:1017): the same
    predictors in the same construction order as AppShadingNetwork with light_pos_freq 8, refrac_freq 2, light_exp_max 5 and
    the refraction-light head clamped at -0.2 (field.py:1373)."""
    default_cfg = {
        "human_light": False, "sphere_direction": False, "light_pos_freq": 8, "inner_init": -0.95,
        "roughness_init": 0.0, "metallic_init": 0.0, "light_exp_max": 5.0, "refrac_freq": 2,
    }
    refrac_exp_max = -0.2


class InfOutNetwork(nn.Module):
    """field.py:1020-1036 -- constructed by the stage-1 renderer (ZT:162) but never evaluated on the path."""

    def __init__(self):
        super().__init__()
        run_dim, d = 256, pe_dim(10)
        mods = []
        for a, b in [(d, run_dim), (run_dim, run_dim), (run_dim, run_dim), (run_dim, run_dim), (run_dim, 3)]:
            mods += [WNLinear(nn.Linear(a, b)), _Slot()]
        self.module0 = nn.Sequential(*mods)


class IoRNetwork(nn.Module):
    """field.py:1046-1059 -- PE-6 -> 256 -> 256 -> 256 -> 1, sigmoid.  Sequential indices 0,2,4,5."""

    def __init__(self):
        super().__init__()
        run_dim, d = 256, pe_dim(6)
        self.module0 = nn.Sequential(
            WNLinear(nn.Linear(d, run_dim)), _Slot(), WNLinear(nn.Linear(run_dim, run_dim)), _Slot(),
            WNLinear(nn.Linear(run_dim, run_dim)), WNLinear(nn.Linear(run_dim, 1)), _Slot())


class ThicknessNetwork(nn.Module):
    """field.py:1068-1081 -- same layout as IoRNetwork; constructed by Stage2Renderer (ZT:951) and only evaluated by the
    non-zero-thickness renderer (network/renderer.py), kept here so that seeds and checkpoints line up."""

    def __init__(self):
        super().__init__()
        run_dim, d = 256, pe_dim(6)
        self.module0 = nn.Sequential(
            WNLinear(nn.Linear(d, run_dim)), _Slot(), WNLinear(nn.Linear(run_dim, run_dim)), _Slot(),
            WNLinear(nn.Linear(run_dim, run_dim)), WNLinear(nn.Linear(run_dim, 1)), _Slot())


class AppShadingNetwork_S2(nn.Module):
    """field.py:786-803 -- the stage-2 surface shader owns no parameters: it evaluates the stage-1 colour network's
    predictors (registered here again as `stage1_network`, as the reference does, so state_dict keys match)."""
    default_cfg = {"human_light": False, "sphere_direction": True, "light_pos_freq": 6, "inner_init": -0.95,
                   "roughness_init": 0.0, "metallic_init": 0.0, "light_exp_max": 5.0, "refrac_freq": 6}

    def __init__(self, cfg, stage1):
        super().__init__()
        self.cfg = {**self.default_cfg, **cfg}
        if self.cfg["human_light"]:
            raise NotImplementedError("stage-2 shader: the human_light variant is not built")
        if bool(self.cfg["sphere_direction"]) != bool(stage1.color_network.cfg["sphere_direction"]):
            raise ValueError("shader_config.sphere_direction of stage 2 and of the stage-1 network must agree: the stage-2 "
                             "shader evaluates the stage-1 outer_light (72 or 144 inputs, field.py:594-597, :829-833)")
        self.stage1_network = stage1

"""ctypes binding of libnunerf_b200.so (the C-ABI declared in include/nunerf.h).

There is no fallback: if the shared library is missing the import raises, and every entry point raises
RuntimeError (with nunerf_last_error()) on a non-zero return code.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NUNERF_LIB", os.path.join(_HERE, "libnunerf_b200.so"))     # NUNERF_LIB: A/B builds of the library

if not os.path.exists(LIB_PATH):
    raise ImportError(f"{LIB_PATH} not found: build it with `make` (or __graft_entry__.build()); "
                      "nu_nerf_b200 has no CPU / PyTorch fallback")
lib = C.CDLL(LIB_PATH)

vp, ci, cf, cll = C.c_void_p, C.c_int, C.c_float, C.c_longlong


class LinearT(C.Structure):
    _fields_ = [("A", vp), ("lda", ci), ("a_lo_off", ci), ("B", vp), ("ldb", ci), ("b_lo_off", ci),
                ("M", ci), ("N", ci), ("K", ci), ("bias", vp), ("act", ci),
                ("aux", vp), ("ldaux", ci), ("aux_lo_off", ci), ("aux_mode", ci),
                ("add", vp), ("ldadd", ci), ("add_lo_off", ci),
                ("mask_in", vp), ("ldmask_in", ci), ("mask_out", vp), ("ldmask_out", ci), ("out_scale", cf),
                ("out", vp), ("ldo", ci), ("out_lo_off", ci), ("out_f32", vp), ("ldo32", ci),
                ("n_store", ci), ("impl", ci)]


class DwT(C.Structure):
    _fields_ = [("dZ", vp), ("ldz", ci), ("z_lo_off", ci), ("X", vp), ("ldx", ci), ("x_lo_off", ci),
                ("M", ci), ("N", ci), ("K", ci), ("dW", vp), ("lddw", ci), ("impl", ci), ("db", vp)]


class SdfInferT(C.Structure):
    _fields_ = [("pts", vp), ("M", ci), ("w", vp * 9), ("ldw", ci * 9), ("bias", vp * 9), ("sdf", vp), ("ld_sdf", ci), ("timeline", vp)]


class ChainLayerT(C.Structure):
    _fields_ = [("w", vp), ("ldw", ci), ("N", ci), ("K", ci), ("n_real", ci), ("bias", vp), ("act", ci),
                ("mask_out", vp), ("ldmask_out", ci), ("mask_in", vp), ("ldmask_in", ci), ("store", vp), ("ld_store", ci),
                ("out32", vp), ("ldo32", ci), ("n32", ci), ("keep", ci), ("cat_pe", ci), ("aux_mode", ci),
                ("aux1", vp), ("ld_aux1", ci), ("aux2", vp), ("ld_aux2", ci), ("e_out", vp), ("ld_e", ci), ("mask_perm", ci)]


class MlpChainT(C.Structure):
    _fields_ = [("x", vp), ("ldx", ci), ("K0", ci), ("pts", vp), ("M", ci), ("n_layers", ci), ("layer", ChainLayerT * 10), ("timeline", vp)]


class SdfAlphaT(C.Structure):
    _fields_ = [("M", ci), ("cos_anneal", cf), ("inv_s_dev", vp), ("sdf", vp), ("ld_sdf", ci), ("grad", vp),
                ("dists", vp), ("dirs", vp), ("alpha", vp), ("grad_err", vp), ("d_alpha", vp), ("d_grad_err", vp),
                ("d_sdf", vp), ("d_grad", vp), ("d_inv_s", vp), ("d_dists", vp), ("d_dirs", vp)]


class ShadeEncodeT(C.Structure):
    _fields_ = [("M", ci), ("pts", vp), ("grad", vp), ("dirs", vp), ("rough_raw", vp), ("ld_rough", ci),
                ("x_outer", vp), ("ld_outer", ci), ("lo_outer", ci), ("x_inner", vp), ("ld_inner", ci), ("lo_inner", ci),
                ("x_weight", vp), ("ld_weight", ci), ("lo_weight", ci), ("x_refrac", vp), ("ld_refrac", ci),
                ("lo_refrac", ci), ("nov", vp), ("d_x_outer", vp), ("ld_dxo", ci), ("d_x_inner", vp), ("ld_dxi", ci),
                ("d_nov", vp), ("d_grad", vp), ("d_rough_raw", vp), ("ld_drough", ci), ("refl", vp),
                ("d_x_refrac", vp), ("ld_dxr", ci), ("d_pts", vp), ("d_dirs", vp), ("pos_freq", ci), ("refrac_freq", ci), ("sphere_direction", ci)]


class ShadeMixT(C.Structure):
    _fields_ = [("M", ci), ("exp_max", cf), ("metallic", vp), ("rough", vp), ("albedo", vp), ("trans", vp),
                ("ld_mat", ci), ("outer", vp), ("ld_outer", ci), ("inner", vp), ("ld_inner", ci), ("weight", vp),
                ("ld_weight", ci), ("refrac", vp), ("ld_refrac", ci), ("nov", vp), ("lut", vp), ("color", vp),
                ("trans_out", vp), ("metallic_out", vp), ("occ_prob", vp), ("d_color", vp), ("d_trans_out", vp),
                ("d_metallic_out", vp), ("dz_metallic", vp), ("dz_albedo", vp), ("dz_trans", vp), ("dz_outer", vp),
                ("dz_inner", vp), ("dz_weight", vp), ("dz_refrac", vp), ("ld_dz", ci), ("lo_dz", ci),
                ("d_rough_raw", vp), ("d_nov", vp), ("d_occ_prob", vp), ("exp_max_refrac", cf), ("use_exp_max_refrac", ci)]


class WDesc(C.Structure):
    _fields_ = [("v", vp), ("g", vp), ("N", ci), ("K", ci), ("ld", ci), ("scale", cf),
                ("row_rot", ci), ("col_rot", ci), ("src_row0", ci), ("n_rows", ci),
                ("wk", vp), ("wk_ld", ci), ("wk_lo", ci), ("wk_row_off", ci),
                ("wtk", vp), ("wtk_ld", ci), ("wtk_lo", ci), ("wtk_col_off", ci),
                ("inv_norm", vp), ("row_f32", vp), ("bias_src", vp), ("bias_dst", vp),
                ("dW", vp), ("lddw", ci), ("dw_row_off", ci), ("dv", vp), ("dg", vp), ("db", vp), ("dbias", vp)]


class BvhNode(C.Structure):
    _fields_ = [("lo", cf * 12), ("hi", cf * 12), ("child", ci * 4), ("count", ci * 4)]


lib.nunerf_last_error.restype = C.c_char_p
lib.nunerf_launch_count.restype = cll
lib.nunerf_mc_blocks.argtypes = [ci]
lib.nunerf_mc_blocks.restype = cll

_SIGS = {
    "nunerf_linear": [C.POINTER(LinearT), vp],
    "nunerf_linear_dw": [C.POINTER(DwT), vp],
    "nunerf_sdf_infer": [C.POINTER(SdfInferT), vp],
    "nunerf_mlp_chain": [C.POINTER(MlpChainT), vp],
    "nunerf_colsum": [vp, ci, ci, ci, ci, vp, vp],
    "nunerf_to_planes": [vp, ci, ci, ci, ci, cf, vp, ci, ci, ci, ci, ci, ci, vp],
    "nunerf_from_planes": [vp, ci, ci, ci, ci, vp, ci, vp],
    "nunerf_f32_to_planes": [vp, ci, vp, ci, ci, ci, ci, vp, ci, ci, ci, vp],
    "nunerf_ray_setup": [vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, vp, vp, vp],
    "nunerf_points": [vp, vp, vp, ci, ci, vp, vp],
    "nunerf_points_bwd": [vp, vp, ci, ci, vp, vp, vp],
    "nunerf_upsample": [vp, vp, vp, vp, ci, ci, ci, vp, cf, vp, vp, vp, vp, vp, vp],
    "nunerf_merge_sdf": [vp, vp, vp, ci, ci, ci, vp, vp],
    "nunerf_probe_weights": [vp, vp, ci, ci, vp, ci, vp, vp, vp, vp],
    "nunerf_render_geometry": [vp, vp, vp, ci, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp],
    "nunerf_inner_counts": [vp, vp, vp, ci, ci, vp, vp],
    "nunerf_seg_composite_fwd": [vp, vp, ci, ci, vp, vp, vp],
    "nunerf_seg_composite_bwd": [vp, vp, ci, ci, vp, vp, vp, vp, vp],
    "nunerf_alpha_importance": [vp, vp, ci, ci, ci, vp, vp, vp],
    "nunerf_composite_fwd": [vp, vp, vp, vp, vp, ci, ci, ci, vp, vp, vp, vp, vp, vp, vp],
    "nunerf_composite_bwd": [vp, vp, vp, vp, vp, ci, ci, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp],
    "nunerf_scatter_rows": [vp, ci, ci, vp, vp, vp],
    "nunerf_encode_pe": [vp, ci, ci, ci, vp, ci, ci, ci, ci, ci, vp],
    "nunerf_sdf_grad_pe": [vp, vp, ci, vp, ci, ci, vp, vp],
    "nunerf_sdf_grad_pe_bwd": [vp, vp, ci, vp, ci, ci, ci, ci, vp, ci, ci, ci, ci, vp],
    "nunerf_sdf_alpha_fwd": [C.POINTER(SdfAlphaT), vp],
    "nunerf_sdf_alpha_bwd": [C.POINTER(SdfAlphaT), vp],
    "nunerf_nerf_prep": [vp, vp, ci, vp, vp, vp],
    "nunerf_nerf_out_fwd": [vp, ci, vp, ci, vp, ci, vp, vp, vp],
    "nunerf_nerf_out_bwd": [vp, ci, vp, ci, vp, ci, vp, vp, vp, ci, ci, ci, vp, ci, ci, ci, vp],
    "nunerf_nerf_out_bwd_geo": [vp, ci, vp, ci, vp, ci, vp, vp, vp, ci, ci, ci, vp, ci, ci, ci, vp, vp],
    "nunerf_pe_bwd": [vp, ci, ci, vp, ci, vp, ci, ci, vp, ci, vp],
    "nunerf_sdf_pe_hess": [vp, vp, ci, vp, ci, vp, ci, vp, vp],
    "nunerf_nerf_prep_bwd": [vp, vp, vp, ci, vp, vp, vp],
    "nunerf_shade_encode_fwd": [C.POINTER(ShadeEncodeT), vp],
    "nunerf_shade_encode_bwd": [C.POINTER(ShadeEncodeT), vp],
    "nunerf_ide_encode": [vp, ci, cf, vp, ci, ci, ci, vp],
    "nunerf_shade_mix_fwd": [C.POINTER(ShadeMixT), vp],
    "nunerf_shade_mix_bwd": [C.POINTER(ShadeMixT), vp],
    "nunerf_rowvec_mask": [vp, vp, ci, ci, ci, ci, vp, ci, ci, vp],
    "nunerf_sdf_skip_split": [vp, vp, ci, ci, ci, vp, ci, ci, vp, vp],
    "nunerf_sdf_bwd2_ew": [vp, ci, ci, vp, ci, ci, vp, ci, ci, ci, ci, ci, vp, ci, ci, vp, ci, ci, vp],
    "nunerf_weights_prepare": [vp, vp, vp, ci, vp],
    "nunerf_weights_backward": [vp, vp, vp, ci, vp],
    "nunerf_adam": [vp, vp, vp, vp, cll, cf, cf, cf, cf, ci, vp],
    "nunerf_bvh_build_host": [vp, ci, vp, ci, vp, ci, vp],
    "nunerf_bvh_trace": [vp, vp, vp, vp, vp, ci, cf, vp, vp, vp, vp],
    "nunerf_trace_brute": [vp, ci, vp, vp, ci, cf, vp, vp, vp, vp],
    "nunerf_hit_interp": [vp, vp, vp, vp, vp, ci, vp, vp, vp, vp],
    "nunerf_refract_bounce": [vp, vp, vp, vp, vp, ci, ci, vp, vp, vp, vp],
    "nunerf_hit_interp_bwd": [vp, vp, vp, vp, vp, ci, ci, vp, vp, vp, vp, vp],
    "nunerf_refract_bounce_bwd": [vp, vp, vp, ci, vp, vp, vp, vp, vp, vp, vp],
    "nunerf_shell_bounce": [vp, vp, vp, vp, vp, vp, ci, ci, vp, vp, vp, vp, vp, vp, vp],
    "nunerf_shell_bounce_bwd": [vp, vp, vp, vp, vp, vp, vp, ci, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp],
    "nunerf_grid_points": [ci, cll, ci, vp, vp, vp],
    "nunerf_grid_mask": [vp, vp, ci, ci, cf, vp, vp],
    "nunerf_mc_count": [vp, ci, cf, vp, vp, vp],
    "nunerf_mc_emit": [vp, ci, cf, vp, ci, vp, vp, vp, vp, vp, vp, vp],
    "nunerf_mma_probe": [ci, ci, ci, ci, ci, ci, vp, vp],
}
for _name, _args in _SIGS.items():
    _fn = getattr(lib, _name)
    _fn.argtypes = _args
    _fn.restype = ci

lib.nunerf_bvh_overflow_count.argtypes = [C.POINTER(C.c_uint)]
lib.nunerf_bvh_overflow_count.restype = ci

ALL_SYMBOLS = ["nunerf_last_error", "nunerf_version", "nunerf_launch_count", "nunerf_mc_blocks",
               "nunerf_bvh_overflow_count"] + list(_SIGS)


def ptr(t):
    """Device (or host) pointer of a tensor, None -> NULL."""
    if t is None:
        return None
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def on_device(t):
    """Context for launching on the device (and current stream of the device) that owns tensor `t`: the entry points
    take raw pointers, so a launch from a process whose current device differs from the tensors' would run on the wrong
    GPU.  One process per GPU (torchrun + torch.cuda.set_device(LOCAL_RANK)) never needs it; the renderers use it once
    per call."""
    return torch.cuda.device(t.device)


def require_current_device(t):
    if t.is_cuda and t.device.index != torch.cuda.current_device():
        raise RuntimeError(f"nu_nerf_b200: tensor on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()} "
                           "-- call torch.cuda.set_device() first (one process per GPU) or wrap the call in torch.cuda.device()")


def check(rc, name):
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {lib.nunerf_last_error().decode()}")


def call(name, *args):
    check(getattr(lib, name)(*args, stream()), name)


def launch_count():
    return int(lib.nunerf_launch_count())

"""ctypes binding of the C ABI declared in include/viewfusion_b200.h.

There is exactly one implementation behind these symbols — the sm_100a CUDA library built in-tree by
`view_fusion_b200.build`.  If it is missing, loading fails loudly; there is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VF_B200_LIB") or os.path.join(_HERE, "libviewfusion_b200.so")   # override: A/B builds in profiling scripts

VF_F32, VF_BF16 = 0, 1
VF_MAX_LEVELS = 8

# every symbol include/viewfusion_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "vf_last_error", "vf_abi_version", "vf_device_check",
    "vf_unet_create", "vf_unet_destroy", "vf_unet_num_params", "vf_unet_param_info", "vf_unet_emb_channels",
    "vf_unet_packed_bytes", "vf_unet_pack_weights", "vf_unet_workspace_bytes", "vf_unet_k0", "vf_unet_forward",
    "vf_unet_last_launches", "vf_unet_read_tap", "vf_unet_set_profiling", "vf_unet_profile_read", "vf_unet_profile_launches", "vf_unet_set_stash", "vf_unet_set_capacity", "vf_unet_forward_generation",
    "vf_pack_views", "vf_pack_nchw", "vf_nhwc_to_nchw", "vf_q_sample",
    "vf_compose_ddpm_step", "vf_compose_mse", "vf_step_prepare", "vf_p_sample_step", "vf_unet_act_dtype",
    "vf_embed", "vf_gn_stats", "vf_gn_apply", "vf_upsample2x", "vf_flat_to_padded", "vf_padded_to_flat", "vf_zero_padding", "vf_conv2d", "vf_debug_force_simt", "vf_debug_flags", "vf_debug_counters", "vf_attention",
    "vf_pack_conv_weight",
    "vf_unet_packed_t_bytes", "vf_unet_pack_weights_t", "vf_unet_backward_workspace_bytes", "vf_unet_backward", "vf_unet_backward_plan", "vf_unet_backward_phase",
    "vf_conv2d_wgrad", "vf_unpack_conv_wgrad", "vf_pack_conv_weight_t", "vf_gn_backward", "vf_attention_backward",
    "vf_upsample2x_backward", "vf_zero_insert2x", "vf_add_inplace", "vf_grad8_to_act", "vf_colsum_bias", "vf_embed_backward",
    "vf_adam_chunk_elems", "vf_adam_step", "vf_eval_metrics", "vf_prepare_batch_u8", "vf_debug_gn_splits", "vf_debug_gn_bwd_splits", "vf_debug_gn_bwd_slab", "vf_debug_conv_tiling",
]


class UnetConfig(C.Structure):
    _fields_ = [
        ("in_channel", C.c_int), ("out_channel", C.c_int), ("inner_channel", C.c_int), ("norm_groups", C.c_int),
        ("n_mults", C.c_int), ("channel_mults", C.c_int * VF_MAX_LEVELS),
        ("n_attn_res", C.c_int), ("attn_res", C.c_int * VF_MAX_LEVELS),
        ("res_blocks", C.c_int), ("image_size", C.c_int),
    ]


class Schedule(C.Structure):
    _fields_ = [
        ("gammas", C.c_void_p), ("sqrt_recip_gammas", C.c_void_p), ("sqrt_recipm1_gammas", C.c_void_p),
        ("posterior_log_variance_clipped", C.c_void_p), ("posterior_mean_coef1", C.c_void_p),
        ("posterior_mean_coef2", C.c_void_p), ("num_timesteps", C.c_int),
    ]


class ComposeArgs(C.Structure):
    _fields_ = [
        ("unet_out", C.c_void_p), ("view_offset", C.c_void_p), ("t", C.c_void_p), ("y_t", C.c_void_p),
        ("y_prev", C.c_void_p), ("z", C.c_void_p), ("seed", C.c_uint64), ("offset", C.c_uint64),
        ("add_noise", C.c_int), ("clip_denoised", C.c_int), ("weighting", C.c_int),
        ("B", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("eps_out", C.c_void_p), ("weights_out", C.c_void_p), ("max_v", C.c_int), ("logits_out", C.c_void_p),
        ("step", C.c_void_p),
    ]


class GnShift(C.Structure):
    _fields_ = [("bias", C.c_void_p), ("emb", C.c_void_p), ("img_row", C.c_void_p), ("emb_ld", C.c_int)]


class SampleStepArgs(C.Structure):
    _fields_ = [
        ("packed", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("y_cond", C.c_void_p), ("B", C.c_int), ("n_max", C.c_int), ("cond_channels", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("images", C.c_int), ("view_offset", C.c_void_p), ("angle", C.c_void_p), ("y_t", C.c_void_p), ("y_prev", C.c_void_p),
        ("t_state", C.c_void_p), ("advance", C.c_int), ("noise_ctr", C.c_void_p), ("seed", C.c_uint64), ("z", C.c_void_p),
        ("add_noise", C.c_int), ("clip_denoised", C.c_int), ("weighting", C.c_int),
        ("x0", C.c_void_p), ("img_sample", C.c_void_p), ("unet_out", C.c_void_p),
        ("level", C.c_void_p), ("t_cur", C.c_void_p), ("rec", C.c_void_p),
        ("eps_out", C.c_void_p), ("weights_out", C.c_void_p), ("logits_out", C.c_void_p), ("max_v", C.c_int),
        ("sched", Schedule),
    ]


class ConvArgs(C.Structure):
    _fields_ = [
        ("dtype", C.c_int), ("images", C.c_int), ("H", C.c_int), ("W", C.c_int), ("in_padded", C.c_int),
        ("out_padded", C.c_int), ("n_seg", C.c_int),
        ("src", C.c_void_p * 3), ("src_c", C.c_int * 3), ("ksize", C.c_int * 3), ("stride", C.c_int),
        ("weight", C.c_void_p), ("cout", C.c_int), ("cout_pad", C.c_int), ("bias", C.c_void_p),
        ("emb", C.c_void_p), ("img_row", C.c_void_p), ("emb_ld", C.c_int), ("residual", C.c_void_p),
        ("out", C.c_void_p), ("out_dtype", C.c_int), ("out_ld", C.c_int), ("qkv_split", C.c_int),
        ("out_vt", C.c_void_p), ("stats", C.c_void_p),
    ]


_lib = None


def load() -> C.CDLL:
    """dlopen the in-tree library (no GPU needed for this) and declare the signatures."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m view_fusion_b200.build` "
            "(view_fusion_b200 has no CPU / PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    p, i, f, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    sig = {
        "vf_last_error": (C.c_char_p, []),
        "vf_abi_version": (i, []),
        "vf_device_check": (i, []),
        "vf_unet_create": (i, [C.POINTER(UnetConfig), i, C.POINTER(p)]),
        "vf_unet_destroy": (None, [p]),
        "vf_unet_num_params": (i, [p]),
        "vf_unet_param_info": (i, [p, i, C.c_char_p, i, C.POINTER(C.c_int64), C.POINTER(i)]),
        "vf_unet_emb_channels": (i, [p]),
        "vf_unet_packed_bytes": (sz, [p]),
        "vf_unet_pack_weights": (i, [p, C.POINTER(p), p, p]),
        "vf_unet_workspace_bytes": (sz, [p, i]),
        "vf_unet_k0": (i, [p]),
        "vf_unet_forward": (i, [p, p, p, sz, i, p, p, p, i, p, p, p]),
        "vf_unet_last_launches": (i, [p]),
        "vf_unet_set_profiling": (i, [p, i]),
        "vf_unet_set_stash": (i, [p, i]),
        "vf_unet_set_capacity": (i, [p, i]),
        "vf_unet_forward_generation": (C.c_ulonglong, [p]),
        "vf_unet_profile_read": (i, [p, C.POINTER(C.c_float), C.POINTER(i)]),
        "vf_eval_metrics": (i, [p, p, i, i, i, i, p, p, p]),
        "vf_prepare_batch_u8": (i, [p, p, i, i, i, i, i, p, p, p, p]),
        "vf_debug_gn_splits": (i, [i, i, i, i, i, C.POINTER(i)]),
        "vf_debug_gn_bwd_splits": (i, [i, i, i, i, i]),
        "vf_debug_gn_bwd_slab": (i, [i, i, i, i, i, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "vf_debug_conv_tiling": (i, [p, C.POINTER(i)]),
        "vf_unet_profile_launches": (i, [p, C.POINTER(C.c_float), C.POINTER(i), C.POINTER(i), i]),
        "vf_unet_read_tap": (i, [p, p, C.c_char_p, p, C.POINTER(C.c_int64), p]),
        "vf_pack_views": (i, [p, p, p, i, i, i, i, i, i, i, i, p, p, p]),
        "vf_pack_nchw": (i, [p, i, i, i, i, i, i, p, p]),
        "vf_nhwc_to_nchw": (i, [p, i, i, i, i, i, p, p]),
        "vf_q_sample": (i, [p, p, p, i, i, p, p]),
        "vf_compose_ddpm_step": (i, [C.POINTER(ComposeArgs), C.POINTER(Schedule), p]),
        "vf_compose_mse": (i, [p, p, p, i, i, i, i, p, p, p, f, p]),
        "vf_step_prepare": (i, [p, i, p, i, i, p, p, p, p, p]),
        "vf_p_sample_step": (i, [p, C.POINTER(SampleStepArgs), p]),
        "vf_unet_act_dtype": (i, [p]),
        "vf_embed": (i, [p, p, i, i, p, p, p, p, p, p, i, p, p]),
        "vf_gn_stats": (i, [p, i, p, i, i, i, i, i, p, p]),
        "vf_gn_apply": (i, [p, i, p, i, p, i, p, i, i, i, i, i, i, p, p, i, p, C.POINTER(GnShift), p]),
        "vf_upsample2x": (i, [p, i, i, i, i, i, p, p]),
        "vf_zero_padding": (i, [p, i, i, i, i, i, p]),
        "vf_flat_to_padded": (i, [p, i, i, i, i, i, p, p]),
        "vf_padded_to_flat": (i, [p, i, i, i, i, i, p, p]),
        "vf_conv2d": (i, [C.POINTER(ConvArgs), p]),
        "vf_debug_force_simt": (None, [i]),
        "vf_debug_flags": (None, [i]),
        "vf_debug_counters": (None, [p]),
        "vf_attention": (i, [p, p, i, i, i, i, p, p, p]),
        "vf_pack_conv_weight": (i, [p, i, i, i, i, p, i, i, i, p]),
        "vf_unet_packed_t_bytes": (sz, [p]),
        "vf_unet_pack_weights_t": (i, [p, p, p]),
        "vf_unet_backward_workspace_bytes": (sz, [p]),
        "vf_unet_backward": (i, [p, p, p, sz, p, C.POINTER(p), p]),
        "vf_unet_backward_plan": (i, [p, i, C.POINTER(i)]),
        "vf_unet_backward_phase": (i, [p, p, p, sz, p, C.POINTER(p), i, i, p]),
        "vf_conv2d_wgrad": (i, [C.POINTER(ConvArgs), p, i, p, p]),
        "vf_unpack_conv_wgrad": (i, [p, i, i, i, i, i, p, i, i, p]),
        "vf_pack_conv_weight_t": (i, [p, i, i, i, i, p, i, i, i, i, p]),
        "vf_gn_backward": (i, [p, i, p, i, p, i, p, i, i, i, i, i, i, p, p, i, p, p, p, p, p, i, p, i, C.POINTER(GnShift), p]),
        "vf_attention_backward": (i, [p, p, p, p, p, i, i, i, i, p, p, p]),
        "vf_upsample2x_backward": (i, [p, i, i, i, i, i, p, i, p]),
        "vf_zero_insert2x": (i, [p, i, i, i, i, i, p, p]),
        "vf_add_inplace": (i, [p, p, i, sz, p]),
        "vf_grad8_to_act": (i, [p, sz, i, i, p, p]),
        "vf_colsum_bias": (i, [p, i, i, i, i, i, p, p, p, p, i, i, p]),
        "vf_embed_backward": (i, [p, p, i, i, p, p, p, p, p, i, p, p, p, p, p, p, p, p, p]),
        "vf_adam_chunk_elems": (i, []),
        "vf_adam_step": (i, [p, p, p, i, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, i, p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.vf_abi_version() != 2:
        raise RuntimeError("libviewfusion_b200.so: ABI version mismatch")
    _lib = lib
    return lib


PROBES_PATH = os.path.join(_HERE, "libviewfusion_b200_probes.so")
PROBE_SYMBOLS = ["vf_debug_umma_shift", "vf_debug_umma_rate", "vf_debug_umma_mn", "vf_debug_mufu_rate"]      # include/viewfusion_b200_probes.h
_probes = None


def load_probes() -> C.CDLL:
    """The probe library (tests / scripts only): product objects + csrc/k_debug.cu.  The product package never calls this."""
    global _probes
    if _probes is None:
        if not os.path.exists(PROBES_PATH):
            raise RuntimeError(f"{PROBES_PATH} is missing: build it with `python -m view_fusion_b200.build`")
        lib = C.CDLL(PROBES_PATH)
        p, i = C.c_void_p, C.c_int
        for name, args in {"vf_debug_mufu_rate": [i, i, i, p, p, p], "vf_debug_umma_rate": [i, i, i, i, i, p, p], "vf_debug_umma_mn": [p, i, p, i, i, i, i, p, p],
                           "vf_debug_umma_shift": [p, i, p, i, i, p, p]}.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = i, args
        lib.vf_last_error.restype = C.c_char_p
        _probes = lib
    return _probes


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().vf_last_error().decode(errors="replace")
        raise RuntimeError(f"viewfusion_b200 {what} failed ({rc}): {msg}")


def ptr(t) -> int:
    """Device pointer of a torch tensor (None -> NULL)."""
    return 0 if t is None else t.data_ptr()


def stream_handle() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


_checked = False


def require_device() -> C.CDLL:
    """The product path: library present AND an sm_100 device.  Raises otherwise."""
    global _checked
    lib = load()
    if not _checked:
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("view_fusion_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        check(lib.vf_device_check(), "device check")
        _checked = True
    return lib

"""ctypes binding of libsin3dm_b200.so (include/sin3dm_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``python -m sin3dm_b200.build``.
There is NO fallback: if the shared library is missing or a call fails, this raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsin3dm_b200.so")

S3D_MAX_LEVELS = 8
S3D_NCOEF = 12
DDPM, DDIM, DDIM_REVERSE = 0, 1, 2
START_X, EPSILON = 0, 1


class UNetConfig(C.Structure):
    _fields_ = [("in_channels", C.c_int), ("model_channels", C.c_int), ("out_channels", C.c_int),
                ("num_res_blocks", C.c_int), ("n_levels", C.c_int), ("channel_mult", C.c_int * S3D_MAX_LEVELS),
                ("use_scale_shift_norm", C.c_int), ("rollout", C.c_int), ("precision", C.c_int),
                ("conv_impl", C.c_int)]


class SchedArgs(C.Structure):
    _fields_ = [("kind", C.c_int), ("mean_type", C.c_int), ("clip_denoised", C.c_int), ("is_mask_t0", C.c_int),
                ("B", C.c_int), ("C", C.c_int), ("n_per_sample", C.c_int64), ("model_out", C.c_void_p), ("x", C.c_void_p),
                ("noise", C.c_void_p), ("y0", C.c_void_p), ("mask", C.c_void_p), ("sample", C.c_void_p),
                ("pred_xstart", C.c_void_p), ("coef_dev", C.c_void_p), ("t_idx_dev", C.c_void_p),
                ("seed", C.c_uint64), ("sample_base", C.c_uint32)]


class VbArgs(C.Structure):
    _fields_ = [("mean_type", C.c_int), ("clip_denoised", C.c_int), ("B", C.c_int), ("n_per_sample", C.c_int64),
                ("x_start", C.c_void_p), ("x_t", C.c_void_p), ("model_out", C.c_void_p), ("noise", C.c_void_p),
                ("pred_xstart", C.c_void_p), ("coef_dev", C.c_void_p), ("logvar_dev", C.c_void_p),
                ("t_idx_dev", C.c_void_p), ("workspace", C.c_void_p), ("out", C.c_void_p)]


class AdamWArgs(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("ema", C.c_void_p * 4), ("ema_rate", C.c_float * 4), ("n_ema", C.c_int), ("n", C.c_int64),
                ("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double), ("weight_decay", C.c_double),
                ("step", C.c_int)]


class LoopArgs(C.Structure):
    _fields_ = [("kind", C.c_int), ("mean_type", C.c_int), ("clip_denoised", C.c_int), ("is_mask_t0", C.c_int),
                ("n_steps", C.c_int), ("t_start", C.c_int), ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("D", C.c_int),
                ("n_per_sample", C.c_int64), ("x_dev", C.c_void_p), ("pred_xstart_dev", C.c_void_p), ("coef_dev", C.c_void_p),
                ("film_dev", C.c_void_p), ("step_noise_dev", C.c_void_p), ("y0_dev", C.c_void_p),
                ("mask_dev", C.c_void_p), ("seed", C.c_uint64), ("sample_base", C.c_uint32), ("use_graph", C.c_int)]


class DecoderConfig(C.Structure):
    _fields_ = [("geo_feat_channels", C.c_int), ("tex_feat_channels", C.c_int), ("feat_channel_up", C.c_int),
                ("mlp_hidden_channels", C.c_int), ("mlp_hidden_layers", C.c_int), ("use_tex", C.c_int),
                ("tex_channels", C.c_int), ("ks", C.c_int), ("precision", C.c_int), ("mlp_impl", C.c_int),
                ("mlp_kind", C.c_int), ("net_kind", C.c_int)]


# every symbol include/sin3dm_b200.h declares: (restype, argtypes)
_SIGNATURES = {
    "s3d_abi_version": (C.c_int, []),
    "s3d_last_error": (C.c_char_p, []),
    "s3d_unet_create": (C.c_int, [C.POINTER(UNetConfig), C.c_int, C.POINTER(C.c_void_p)]),
    "s3d_unet_destroy": (C.c_int, [C.c_void_p]),
    "s3d_unet_num_tensors": (C.c_int, [C.c_void_p]),
    "s3d_unet_tensor_info": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int),
                                       C.POINTER(C.c_int64)]),
    "s3d_unet_load_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int]),
    "s3d_unet_finalize": (C.c_int, [C.c_void_p]),
    "s3d_unet_film_dim": (C.c_int, [C.c_void_p]),
    "s3d_unet_film": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "s3d_unet_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_void_p]),
    "s3d_unet_forward_film": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                        C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "s3d_unet_last_launches": (C.c_int, [C.c_void_p]),
    "s3d_unet_workspace_bytes": (C.c_int64, [C.c_void_p]),
    "s3d_unet_op_count": (C.c_int, [C.c_void_p]),
    "s3d_unet_op_info": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_double)]),
    "s3d_unet_profile_ops": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "s3d_unet_profile_mode": (C.c_int, [C.c_void_p]),
    "s3d_unet_trace_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "s3d_trace_slots": (C.c_int, []),
    "s3d_unet_trace_read": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "s3d_sched_step": (C.c_int, [C.POINTER(SchedArgs), C.c_void_p]),
    "s3d_q_sample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64,
                               C.c_void_p]),
    "s3d_philox_normal": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_uint64, C.c_uint32, C.c_uint32,
                                    C.c_void_p]),
    "s3d_vb_workspace_bytes": (C.c_int64, [C.c_int, C.c_int64]),
    "s3d_vb_terms": (C.c_int, [C.POINTER(VbArgs), C.c_void_p]),
    "s3d_plane_mse": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                C.c_void_p]),
    "s3d_adamw_ema_step": (C.c_int, [C.POINTER(AdamWArgs), C.c_void_p]),
    "s3d_sample_loop": (C.c_int, [C.c_void_p, C.POINTER(LoopArgs), C.c_void_p]),
    "s3d_unet_graph_builds": (C.c_int, [C.c_void_p]),
    "s3d_unet_set_training": (C.c_int, [C.c_void_p, C.c_int]),
    "s3d_unet_refresh_dev": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_void_p]),
    "s3d_unet_grad_numel": (C.c_int64, [C.c_void_p]),
    "s3d_unet_grad_offset": (C.c_int64, [C.c_void_p, C.c_int]),
    "s3d_unet_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "s3d_unet_bwd_op_count": (C.c_int, [C.c_void_p]),
    "s3d_unet_bwd_op_info": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_double)]),
    "s3d_unet_profile_bwd_ops": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "s3d_decoder_create": (C.c_int, [C.POINTER(DecoderConfig), C.c_int, C.POINTER(C.c_void_p)]),
    "s3d_decoder_destroy": (C.c_int, [C.c_void_p]),
    "s3d_decoder_num_tensors": (C.c_int, [C.c_void_p]),
    "s3d_decoder_tensor_info": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int),
                                          C.POINTER(C.c_int64)]),
    "s3d_decoder_load_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int]),
    "s3d_decoder_finalize": (C.c_int, [C.c_void_p]),
    "s3d_decoder_set_planes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p]),
    "s3d_decoder_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_float), C.c_int, C.c_void_p,
                                     C.c_void_p]),
    "s3d_decoder_decode_grid": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                          C.POINTER(C.c_float), C.c_int, C.c_void_p, C.c_void_p]),
    "s3d_decoder_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p]),
    "s3d_decoder_last_launches": (C.c_int, [C.c_void_p]),
    "s3d_decoder_planes_read": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64]),
    "s3d_unet_debug_count": (C.c_int, [C.c_void_p]),
    "s3d_unet_debug_info": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int),
                                      C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "s3d_unet_debug_read": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64]),
}

_lib = None


class S3DError(RuntimeError):
    pass


def lib():
    """The loaded shared library (raises if it has not been built — there is no CPU/torch fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise S3DError(f"{LIB_PATH} is missing: build it with `python -m sin3dm_b200.build` "
                           "(sin3dm_b200 has no fallback path)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(l, name)          # AttributeError here == header / library mismatch
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        raise S3DError(lib().s3d_last_error().decode())


def current_stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)

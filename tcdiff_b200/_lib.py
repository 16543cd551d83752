"""ctypes binding of libtcdiff_sm100a.so (the C-ABI declared in include/tcdiff_b200.h).

There is no fallback of any kind: if the shared library is missing or a call fails, this raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libtcdiff_sm100a.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_GELU, ACT_MISH, ACT_SILU, ACT_LEAKY_RELU = 0, 1, 2, 3, 4, 5

_p, _i, _l, _f, _d, _u = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_double, ctypes.c_uint32

# name -> argtypes (every entry returns int unless listed in _RESTYPES)
SIGNATURES = {
    "tcd_cfg_ddim_step": [_p, _p, _p, _p, _p, _p, _p, _p, _l, _l, _i, _f, _f, _f, _f, _f, _f, _i, _i, _p],
    "tcd_cfg_ddpm_step": [_p, _p, _p, _p, _p, _p, _l, _l, _i, _f, _f, _f, _f, _i, _i, _f, _f, _p, _p, _p],
    "tcd_cfg_ddim_step_rng": [_p, _p, _p, _p, _u, _p, _p, _p, _p, _l, _l, _i, _f, _f, _f, _f, _f, _f, _i, _i, _p],
    "tcd_cfg_ddpm_step_rng": [_p, _p, _p, _p, _u, _p, _p, _l, _l, _i, _f, _f, _f, _f, _i, _i, _f, _f, _p, _p, _p],
    "tcd_philox_normal": [_p, _l, _p, _u, _p],
    "tcd_inpaint_traj": [_p, _p, _p, _l, _l, _i, _p],
    "tcd_q_sample": [_p, _p, _p, _p, _p, _p, _p, _p, _l, _i, _i, _i, _i, _i, _i, _p],
    "tcd_ax_from_6v": [_p, _p, _l, _p],
    "tcd_smpl_fk": [_p, _p, _p, _l, _p],
    "tcd_motion_fk": [_p, _p, _l, _i, _p],
    "tcd_loss_workspace_floats": [_i, _i, _i],
    "tcd_loss_backward": [_p, _p, _p, _f, _p, _i, _i, _i, _i, _p],
    "tcd_loss_forward": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p],
    "tcd_gemm": [_i, _p, _l, _p, _l, _p, _i, _i, _p, _l, _l, _l, _l, _p],
    "tcd_layernorm_rotary": [_i, _p, _p, _p, _f, _p, _p, _p, _p, _l, _i, _i, _p],
    "tcd_rotary": [_p, _p, _p, _p, _l, _i, _i, _p],
    "tcd_film_residual_norm": [_i, _p, _p, _p, _i, _p, _p, _f, _p, _l, _l, _p, _p, _f, _p, _p, _p, _p, _l, _i, _i, _p],
    "tcd_gemm_film_residual_norm": [_p, _l, _p, _l, _p, _l, _l, _p, _p, _p, _p, _f, _p, _l, _l, _p, _p, _f, _p, _p, _p, _p,
                                    _l, _i, _p],
    "tcd_gemm_frn_set_debug": [_p],
    "tcd_attention": [_i, _p, _l, _l, _p, _l, _l, _p, _l, _l, _p, _l, _l, _i, _i, _i, _i, _f, _p],
    "tcd_time_embed": [_i, _p, _p, _p, _i, _i, _i, _p],
    "tcd_cond_pool": [_i, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p],
    "tcd_time_cond": [_i, _p, _p, _p, _p, _p, _p, _i, _i, _p],
    "tcd_sampler_time_cond": [_i, _p, _p, _p, _p, _i, _i, _i, _p],
    "tcd_build_memory": [_i, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p],
    "tcd_scatter_rows": [_i, _p, _l, _l, _p, _l, _l, _l, _i, _i, _i, _p],
    "tcd_masked_blend": [_p, _p, _p, _i, _i, _i, _i, _p],
    "tcd_convert_pad": [_i, _p, _l, _p, _l, _l, _i, _p],
    "tcd_cast_transpose": [_i, _p, _l, _p, _l, _l, _l, _p],
    "tcd_pack_weights": [_p, _i, _i, _p],
    "tcd_group_colsum": [_p, _p, _l, _l, _l, _i, _p, _l, _i, _p],
    "tcd_act_forward": [_i, _p, _p, _l, _p],
    "tcd_act_backward": [_i, _p, _p, _p, _l, _p],
    "tcd_layernorm_backward_partials": [_l],
    "tcd_layernorm_backward": [_p, _p, _p, _f, _p, _p, _p, _l, _i, _p],
    "tcd_film_backward": [_p, _p, _p, _l, _l, _p, _p, _l, _l, _i, _i, _i, _p],
    "tcd_attention_backward_workspace_floats": [_i, _i, _i],
    "tcd_attention_backward": [_p, _l, _l, _p, _l, _l, _p, _l, _l, _p, _l, _l, _p, _l, _l, _p, _l, _l, _p, _l, _l, _p, _l, _l,
                               _p, _i, _i, _i, _i, _f, _p],
    "tcd_act_forward_bf16": [_i, _p, _p, _l, _f, _p, _u, _p],
    "tcd_act_backward_bf16": [_i, _p, _p, _p, _l, _f, _p, _u, _p],
    "tcd_layernorm_backward_mixed": [_i, _i, _p, _p, _p, _p, _p, _p, _i, _f, _p, _p, _p, _p, _l, _i, _f, _p, _u, _p],
    "tcd_layernorm_backward_mixed_partials": [_l],
    "tcd_layernorm_bf16": [_p, _p, _p, _f, _p, _l, _i, _f, _p, _u, _p],
    "tcd_film_residual_bf16": [_p, _p, _p, _l, _l, _p, _l, _i, _i, _f, _p, _u, _p],
    "tcd_film_backward_workspace_floats": [_i, _i, _i],
    "tcd_film_backward_bf16": [_p, _p, _p, _l, _l, _p, _p, _l, _l, _p, _i, _i, _i, _f, _p, _u, _p],
    "tcd_colsum_bf16_workspace_floats": [_l, _i],
    "tcd_colsum_bf16": [_p, _l, _l, _i, _p, _p, _p],
    "tcd_gemm_tn_workspace_floats": [_l, _l, _l],
    "tcd_gemm_tn": [_p, _l, _p, _l, _p, _l, _l, _l, _l, _p, _p],
    "tcd_attention_train_workspace_floats": [_i, _i, _i],
    "tcd_attention_train_forward": [_p, _l, _l, _p, _l, _l, _p, _l, _l, _p, _l, _l, _p, _i, _i, _i, _i, _f, _f, _p, _u, _p],
    "tcd_dropout": [_i, _p, _p, _l, _f, _p, _u, _p],
    "tcd_dropout_mask_attention": [_p, _i, _i, _i, _i, _f, _p, _u, _p],
    "tcd_attention_train_backward": [_p, _l, _l, _p, _l, _l, _p, _l, _l, _p, _l, _l, _p, _l, _l, _p, _p, _l, _l, _p, _l, _l,
                                     _p, _l, _l, _p, _i, _i, _i, _i, _f, _f, _p, _u, _p],
    "tcd_samples_to_poses": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p],
    "tcd_samples_to_poses_long": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p],
    "tcd_lstm_layer": [_p, _l, _l, _p, _p, _p, _p, _p, _l, _l, _p, _i, _i, _i, _p],
    "tcd_attention_f32_hd": [_i, _p, _l, _l, _p, _l, _l, _p, _l, _l, _p, _l, _l, _i, _i, _i, _i, _f, _p],
    "tcd_kalman_smooth": [_p, _p, _p, _i, _i, _d, _p],
    "tcd_adan_ema_step": [_p, _p, _p, _p, _p, _p, _p, _l, _l, _d, _d, _d, _d, _d, _d, _d, _d, _p],
    "tcd_adan_ema_step_device": [_p, _p, _p, _p, _p, _p, _p, _l, _p, _p, _d, _d, _d, _d, _d, _d, _d, _d, _p],
    "tcd_ema_update": [_p, _p, _l, _d, _p],
    "tcd_ema_update_multi": [_p, _p, _p, _i, _l, _d, _p],
    "tcd_tuning": [ctypes.c_char_p],
    "tcd_last_error": [],
    "tcd_version": [],
    "tcd_arch": [],
}
_RESTYPES = {"tcd_last_error": ctypes.c_char_p, "tcd_arch": ctypes.c_char_p, "tcd_loss_workspace_floats": _l, "tcd_attention_train_workspace_floats": _l, "tcd_gemm_tn_workspace_floats": _l, "tcd_layernorm_backward_mixed_partials": _l, "tcd_film_backward_workspace_floats": _l, "tcd_colsum_bf16_workspace_floats": _l,
             "tcd_layernorm_backward_partials": _l, "tcd_attention_backward_workspace_floats": _l}

_lib = None


class TcdError(RuntimeError):
    pass


def lib():
    """The loaded library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "tcdiff_b200: %s not found. Build it with `python -m tcdiff_b200.build` "
                "(nvcc, sm_100a). There is no CPU or PyTorch fallback." % LIB_PATH)
        h = ctypes.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(h, name)  # AttributeError if the .so does not export a declared symbol
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, ctypes.c_int)
        _lib = h
    return _lib


LOSS_L2, LOSS_L1 = 0, 1      # include/tcdiff_b200.h TCD_LOSS_*
LAUNCHES = [0]   # number of C-ABI kernel-launching calls made by this process (bench.py reports it)


def check(rc):
    LAUNCHES[0] += 1
    if rc != 0:
        raise TcdError("libtcdiff_sm100a: %s (code %d)" % (lib().tcd_last_error().decode(), rc))

"""ctypes binding of ``libicadv_b200.so`` (the C ABI declared in ``include/icadv.h``).

There is no fallback: if the shared library is missing or a call fails, this raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ICADV_LIB") or os.path.join(_HERE, "libicadv_b200.so")   # ICADV_LIB: developer A/B of two builds

OK = 0
FORM_SCONV, FORM_TCONV = 0, 1
EPI_LINEAR, EPI_GDN_FWD, EPI_IGDN_FWD, EPI_GDN_BWD, EPI_IGDN_BWD = 0, 1, 2, 3, 4
ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_ABS = 0, 1, 2, 3
RED_BLOCKS = 128
PACK_CONV_FWD, PACK_CONV_DGRAD, PACK_CONVT_FWD, PACK_CONVT_DGRAD = 0, 1, 2, 3

_fp = C.c_void_p  # device pointers travel as integers


class ConvDesc(C.Structure):
    _fields_ = [("form", C.c_int), ("ksize", C.c_int), ("stride", C.c_int), ("n_img", C.c_int),
                ("in_h", C.c_int), ("in_w", C.c_int), ("k_ch", C.c_int), ("n_ch", C.c_int),
                ("inp", _fp), ("wpack", _fp), ("bias", _fp), ("out", _fp),
                ("epi", C.c_int), ("act", C.c_int),
                ("gmat", _fp), ("beta", _fp), ("out_scale", _fp), ("y_prev", _fp), ("sc_prev", _fp),
                ("acc_from_in", C.c_int), ("round_out_tf32", C.c_int), ("in_pad4", C.c_int), ("active", _fp), ("n_active", _fp)]


class PerturbState(C.Structure):
    _fields_ = [("sum_d2", _fp), ("loss_i", _fp), ("branch", _fp), ("active", _fp), ("n_active", _fp),
                ("step", _fp), ("lr", _fp), ("step_size", _fp), ("bc2_sqrt", _fp), ("counter", _fp),
                ("cond_handle", C.c_uint64)]


class IcadvError(RuntimeError):
    pass


_lib = None

_SIGS = {
    "icadv_last_error": (C.c_char_p, []),
    "icadv_version": (C.c_int, []),
    "icadv_check_device": (C.c_int, []),
    "icadv_conv_out_hw": (C.c_int, [C.POINTER(ConvDesc), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "icadv_conv_plan_create": (C.c_int, [C.POINTER(ConvDesc), C.POINTER(C.c_void_p)]),
    "icadv_conv_plan_launch": (C.c_int, [C.c_void_p, C.c_void_p]),
    "icadv_conv_plan_destroy": (C.c_int, [C.c_void_p]),
    "icadv_conv_tc": (C.c_int, [C.POINTER(ConvDesc), C.c_void_p]),
    "icadv_conv_tc_supported": (C.c_int, [C.POINTER(ConvDesc)]),
    "icadv_conv_simt": (C.c_int, [C.POINTER(ConvDesc), C.c_void_p]),
    "icadv_conv_wgrad": (C.c_int, [C.POINTER(ConvDesc), _fp, _fp, _fp, C.c_void_p]),
    "icadv_conv_wgrad_ex": (C.c_int, [C.POINTER(ConvDesc), _fp, _fp, _fp, C.c_int, C.c_void_p]),
    "icadv_conv_wgrad_tc_supported": (C.c_int, [C.POINTER(ConvDesc)]),
    "icadv_pack_weight": (C.c_int, [_fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "icadv_pack_weight_rgb": (C.c_int, [_fp, _fp, C.c_int, C.c_int, C.c_void_p]),
    "icadv_pad_rgb4": (C.c_int, [_fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, _fp, _fp, C.c_void_p]),
    "icadv_conv_plan_num_launches": (C.c_int, [C.c_void_p]),
    "icadv_conv_plan_set_debug": (C.c_int, [C.c_void_p, _fp]),
    "icadv_unpack_weight": (C.c_int, [_fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "icadv_nchw_to_nhwc": (C.c_int, [_fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "icadv_nhwc_to_nchw": (C.c_int, [_fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "icadv_clamp01_nhwc_to_nchw": (C.c_int, [_fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "icadv_clamp01_backward_nchw_to_nhwc": (C.c_int, [_fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "icadv_pixel_shuffle": (C.c_int, [_fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "icadv_copy_channels": (C.c_int, [_fp, _fp, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "icadv_gdn_reparam": (C.c_int, [_fp, _fp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p]),
    "icadv_perturb_forward": (C.c_int, [_fp, _fp, _fp, _fp, C.POINTER(PerturbState), C.c_int, C.c_int64, C.c_float,
                                        C.c_float, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double,
                                        C.c_void_p]),
    "icadv_perturb_update_adam": (C.c_int, [_fp, _fp, _fp, _fp, _fp, _fp, C.POINTER(PerturbState), C.c_int, C.c_int64,
                                            C.c_float, C.c_double, C.c_double, C.c_double, C.c_float, C.c_float,
                                            C.c_void_p]),
    "icadv_perturb_forward_roi": (C.c_int, [_fp, _fp, _fp, _fp, C.POINTER(PerturbState), C.c_int, C.c_int64, C.c_float,
                                            C.c_float, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double,
                                            C.c_double, _fp, C.c_int, C.c_void_p]),
    "icadv_perturb_update_adam_roi": (C.c_int, [_fp, _fp, _fp, _fp, _fp, _fp, C.POINTER(PerturbState), C.c_int,
                                                C.c_int64, C.c_float, C.c_double, C.c_double, C.c_double, C.c_float,
                                                C.c_float, _fp, C.c_void_p]),
    "icadv_output_loss_roi": (C.c_int, [_fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int64, C.c_int, C.c_float, _fp, _fp, _fp,
                                        C.c_void_p]),
    "icadv_ifgsm_update": (C.c_int, [_fp, _fp, _fp, C.c_int64, C.c_float, C.c_float, C.c_void_p]),
    "icadv_cw_combine": (C.c_int, [_fp, _fp, _fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int64, C.c_void_p]),
    "icadv_mifgsm_update": (C.c_int, [_fp, _fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int64, C.c_float, C.c_float, C.c_float,
                                      C.c_void_p]),
    "icadv_output_loss": (C.c_int, [_fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int64, C.c_int, C.c_float, _fp, _fp,
                                    C.c_void_p]),
    "icadv_bound_forward": (C.c_int, [_fp, _fp, C.c_int64, C.c_float, C.c_int, C.c_void_p]),
    "icadv_bound_backward": (C.c_int, [_fp, _fp, _fp, C.c_int64, C.c_float, C.c_int, C.c_void_p]),
    "icadv_eb_prepare": (C.c_int, [C.POINTER(_fp), C.POINTER(_fp), C.POINTER(_fp), _fp, C.c_int, C.c_void_p]),
    "icadv_eb_forward": (C.c_int, [_fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int64, C.c_int, C.c_int,
                                   C.c_float, C.c_float, C.c_void_p]),
    "icadv_gc_forward": (C.c_int, [_fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int64, C.c_int, C.c_float,
                                   C.c_float, C.c_float, C.c_void_p]),
    "icadv_eb_backward_workspace_floats": (C.c_int, [C.c_int64, C.c_int]),
    "icadv_eb_backward": (C.c_int, [_fp, _fp, _fp, C.POINTER(_fp), C.POINTER(_fp), _fp, _fp, _fp, C.c_int64, C.c_int,
                                    C.c_float, C.c_void_p]),
    "icadv_gc_backward": (C.c_int, [_fp, _fp, _fp, _fp, _fp, _fp, _fp, C.c_int64, C.c_float, C.c_float, C.c_void_p]),
    "icadv_log_sum": (C.c_int, [_fp, _fp, _fp, C.c_int64, C.c_float, C.c_void_p]),
    "icadv_log_sum_backward": (C.c_int, [_fp, _fp, C.c_int64, C.c_float, _fp, C.c_float, C.c_void_p]),
    "icadv_scaled_diff": (C.c_int, [_fp, _fp, _fp, C.c_int64, _fp, C.c_float, C.c_void_p]),
    "icadv_gdn_param_grad_workspace_floats": (C.c_int, [C.c_int]),
    "icadv_gdn_param_grad": (C.c_int, [_fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, C.c_int64, C.c_int, C.c_int, C.c_float,
                                       C.c_float, C.c_void_p]),
    "icadv_gdn_param_operands": (C.c_int, [_fp, _fp, _fp, _fp, _fp, C.c_int64, C.c_int, C.c_void_p]),
    "icadv_gdn_param_grad_finalize": (C.c_int, [_fp, _fp, _fp, _fp, _fp, _fp, C.c_int, C.c_float, C.c_float, C.c_void_p]),
    "icadv_sumsq": (C.c_int, [_fp, _fp, _fp, C.c_int64, C.c_void_p]),
    "icadv_adam_clip_step": (C.c_int, [_fp, _fp, _fp, _fp, C.c_int64, _fp, C.c_float, C.c_float, C.c_double, C.c_double,
                                       C.c_double, C.c_double, C.c_int, C.c_void_p]),
    "icadv_unary": (C.c_int, [_fp, _fp, _fp, C.c_int64, C.c_int, C.c_void_p]),
    "icadv_act_backward": (C.c_int, [_fp, _fp, _fp, C.c_int64, C.c_int, C.c_void_p]),
    "icadv_ssim_workspace_floats": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "icadv_ssim_level": (C.c_int, [_fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int,
                                   C.c_int, C.c_float, C.c_float, C.c_void_p]),
    "icadv_ssim_level_backward": (C.c_int, [_fp, _fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_float, C.c_float,
                                            C.c_void_p]),
    "icadv_ssim_vg_workspace_floats": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "icadv_ssim_level_value_grad": (C.c_int, [_fp, _fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float),
                                              C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_void_p]),
    "icadv_msssim_coefficients": (C.c_int, [_fp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_float),
                                            C.POINTER(C.c_float), C.c_int, _fp, _fp, _fp, C.c_int, C.c_int, C.c_void_p]),
    "icadv_ssim_combine": (C.c_int, [_fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p]),
    "icadv_avgpool2": (C.c_int, [_fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "icadv_sum_sqdiff": (C.c_int, [_fp, _fp, _fp, _fp, C.c_int, C.c_int64, C.c_void_p]),
    "icadv_split3": (C.c_int, [_fp, _fp, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "icadv_gdn_bwd_operand_split3": (C.c_int, [_fp, _fp, _fp, _fp, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "icadv_sum_slices": (C.c_int, [_fp, _fp, C.c_int64, C.c_int, C.c_void_p]),
    "icadv_gdn_apply": (C.c_int, [_fp, _fp, _fp, _fp, C.c_int64, C.c_int, C.c_void_p]),
    "icadv_gdn_bwd_combine": (C.c_int, [_fp, _fp, _fp, _fp, _fp, C.c_int64, C.c_int, C.c_void_p]),
    "icadv_probe_tf32_peak": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_double), C.c_void_p]),
    "icadv_uniform_noise": (C.c_int, [_fp, C.c_int64, C.c_uint64, C.c_uint64, C.c_float, C.c_float, C.c_void_p]),
    "icadv_graph_if_begin": (C.c_int, [_fp, C.c_void_p, C.c_void_p]),
    "icadv_graph_if_end": (C.c_int, [C.c_void_p]),
    "icadv_graph_if_create": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "icadv_graph_if_begin_handle": (C.c_int, [C.c_uint64, C.c_void_p, C.c_void_p]),
    "icadv_attention_gate": (C.c_int, [_fp, _fp, _fp, _fp, C.c_int64, C.c_void_p]),
    "icadv_attention_gate_backward": (C.c_int, [_fp, _fp, _fp, _fp, _fp, C.c_int64, C.c_void_p]),
    "icadv_pmf_to_quantized_cdf": (C.c_int, [C.POINTER(C.c_float), C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "icadv_rans_lane_capacity": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "icadv_rans_encode": (C.c_int, [_fp, _fp, C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, C.c_int, C.c_int, _fp, C.c_int,
                                    C.c_int, _fp, _fp, _fp, C.c_int, _fp, C.c_void_p]),
    "icadv_rans_decode": (C.c_int, [_fp, C.c_int, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, C.c_int,
                                    C.c_int, _fp, C.c_int, C.c_int, C.c_void_p]),
    "icadv_rans_decode_init": (C.c_int, [_fp, C.c_int, C.c_int, C.c_int, _fp, _fp, C.c_void_p]),
    "icadv_rans_decode_step": (C.c_int, [_fp, C.c_int, _fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, _fp, _fp, _fp,
                                         C.c_int, _fp, C.c_int, C.c_int, C.c_void_p]),
    "icadv_build_indexes": (C.c_int, [_fp, _fp, C.c_int, C.c_float, _fp, C.c_longlong, C.c_void_p]),
}


def exported_symbols():
    return sorted(_SIGS)


def lib():
    """Load (once) and return the shared library; raises if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise IcadvError(f"{LIB_PATH} is not built (run `python __graft_entry__.py` / csrc/build.py); "
                             "there is no fallback path")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc):
    if rc != OK:
        msg = lib().icadv_last_error()
        raise IcadvError(f"icadv error {rc}: {msg.decode() if msg else ''}")


def call(name, *args):
    check(getattr(lib(), name)(*args))

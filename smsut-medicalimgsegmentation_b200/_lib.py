"""ctypes binding of libsmsut_b200.so (the C ABI declared in include/smsut_b200.h).

The library is the product: there is no CPU or PyTorch fallback.  If the shared object is missing the
import fails loudly, and every call checks the integer status and raises with smsut_last_error().
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsmsut_b200.so")

c_void_p, c_int, c_int64, c_float = C.c_void_p, C.c_int32, C.c_int64, C.c_float

TC_CONV, TC_CONVT_FWD, TC_CONVT_DGRAD = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH = 0, 1, 2, 3


class ConvTcArgs(C.Structure):
    _fields_ = [
        ("kind", c_int), ("ksize", c_int), ("n", c_int), ("h", c_int), ("w", c_int), ("nsrc", c_int),
        ("src", c_void_p * 2), ("src_c", c_int * 2), ("src_ld", c_int * 2),
        ("wpack", c_void_p), ("ncols", c_int), ("ncols_pad", c_int),
        ("out0", c_void_p), ("out0_ld", c_int), ("out0_coff", c_int),
        ("out1", c_void_p), ("out1_ld", c_int), ("out1_coff", c_int), ("split", c_int),
        ("bias", c_void_p), ("act", c_int), ("slope", c_float), ("accumulate", c_int), ("out_f32", c_int),
        ("bn", c_int), ("stats", c_void_p),
    ]


class WgradTcArgs(C.Structure):
    _fields_ = [
        ("kind", c_int), ("ksize", c_int), ("n", c_int), ("h", c_int), ("w", c_int),
        ("x", c_void_p), ("x_c", c_int), ("x_ld", c_int),
        ("dy", c_void_p), ("dy_c", c_int), ("dy_ld", c_int),
        ("dw", c_void_p), ("cin_total", c_int), ("ci_off", c_int), ("cout_total", c_int), ("c_valid", c_int),
        ("dw_layout", c_int),
    ]


class ConvDirectArgs(C.Structure):
    _fields_ = [
        ("n", c_int), ("h", c_int), ("w", c_int), ("cin", c_int),
        ("cout", c_int), ("kh", c_int), ("kw", c_int), ("stride", c_int), ("pad", c_int),
        ("ho", c_int), ("wo", c_int),
        ("x", c_void_p), ("x_ld", c_int), ("x_f32", c_int),
        ("wt", c_void_p), ("bias", c_void_p),
        ("y", c_void_p), ("y_ld", c_int), ("y_f32", c_int),
        ("act", c_int), ("slope", c_float), ("accumulate", c_int),
    ]


class PackEntry(C.Structure):
    _fields_ = [
        ("w", c_void_p), ("fprop", c_void_p), ("dgrad", c_void_p),
        ("cout", c_int), ("cin", c_int), ("kh", c_int), ("kw", c_int),
        ("transposed", c_int), ("cout_pad", c_int), ("cin_pad", c_int),
    ]


class UnpackEntry(C.Structure):
    _fields_ = [("scratch", c_void_p), ("grad", c_void_p), ("rows", c_int), ("cols", c_int), ("taps", c_int),
                ("pad", c_int)]


P = c_void_p
_PROTOS = {
    "smsut_conv_tc": [C.POINTER(ConvTcArgs), P],
    "smsut_conv_tc_fuses_stats": [C.POINTER(ConvTcArgs)],
    "smsut_wgrad_tc": [C.POINTER(WgradTcArgs), P],
    "smsut_conv_direct_fprop": [C.POINTER(ConvDirectArgs), P],
    "smsut_conv_direct_dgrad": [C.POINTER(ConvDirectArgs), P],
    "smsut_conv_direct_wgrad": [C.POINTER(ConvDirectArgs), P, P, P],
    "smsut_head1x1_bwd": [P, P, P, P, P, P, P, c_int64, c_int, c_int, P],
    "smsut_in_stats": [P, c_int, c_int, c_int, P, P],
    "smsut_in_apply": [P, P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_float, P],
    "smsut_in_bwd_reduce": [P, P, P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_float, P],
    "smsut_in_bwd_apply": [P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int,
                           c_float, P],
    "smsut_in_bwd_fused": [P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int,
                           c_float, P],
    "smsut_in_bwd2_fused": [P, P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, P],
    "smsut_in_bwd2_reduce": [P, P, P, P, P, c_int, c_int, c_int, P],
    "smsut_in_bwd2_apply": [P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, P],
    "smsut_bn_pool": [P, P, c_int, c_int, c_int, P],
    "smsut_bn_running_update": [P, c_int, c_int, c_int, c_int, c_float, P, P, P],
    "smsut_bn_eval_stats": [P, P, P, c_int, c_int, c_int, c_int, P],
    "smsut_act_fwd": [P, P, c_int64, c_int, c_float, P],
    "smsut_act_bwd": [P, P, P, P, c_int64, c_int, c_float, P],
    "smsut_add_bf16": [P, P, P, c_int64, P],
    "smsut_colsum_bf16": [P, c_int, c_int, P, P],
    "smsut_maxpool2_fwd": [P, P, c_int, c_int, c_int, c_int, P],
    "smsut_maxpool2_bwd": [P, P, P, P, c_int, c_int, c_int, c_int, P],
    "smsut_avgpool2_fwd": [P, P, c_int, c_int, c_int, c_int, P],
    "smsut_avgpool2_bwd": [P, P, P, c_int, c_int, c_int, c_int, P],
    "smsut_bilinear2_fwd": [P, P, c_int, c_int, c_int, c_int, P],
    "smsut_bilinear2_bwd": [P, P, c_int, c_int, c_int, c_int, P],
    "smsut_nchw_f32_to_nhwc_bf16": [P, P, c_int, c_int, c_int, c_int, c_int, P],
    "smsut_nhwc_bf16_to_nchw_f32": [P, P, c_int, c_int, c_int, c_int, c_int, P],
    "smsut_build_tsl_input": [P, P, P, c_int, c_int, c_int, c_int, P],
    "smsut_dice_ce_fwd": [P, P, P, P, c_int64, c_int, P],
    "smsut_dice_ce_finish": [P, P, c_int64, c_int, c_float, c_float, P],
    "smsut_dice_ce_bwd": [P, P, P, P, P, c_float, P, c_int64, c_int64, c_int, c_float, c_float, P],
    "smsut_softmax_mse_fwd": [P, P, P, c_int64, c_int, P],
    "smsut_softmax_mse_bwd": [P, P, P, P, c_int64, c_int, P],
    "smsut_heads_split_fwd": [P, P, c_int64, c_int, c_int, P],
    "smsut_heads_split_bwd": [P, P, c_int64, c_int, c_int, P],
    "smsut_wce_fwd": [P, P, P, P, P, c_int64, c_int, P],
    "smsut_wce_bwd": [P, P, P, P, P, P, c_int, P, c_int64, c_int, P],
    "smsut_softmax_mse_masked_fwd": [P, P, P, c_int, P, c_int64, c_int, P],
    "smsut_softmax_mse_masked_bwd": [P, P, P, c_int, P, P, P, c_int64, c_int, P],
    "smsut_argmax_c": [P, P, c_int64, c_int, P],
    "smsut_confusion_counts": [P, P, P, c_int64, c_int, P],
    "smsut_l1_fwd": [P, P, P, c_int64, c_float, P],
    "smsut_l1_bwd": [P, P, P, c_float, P, c_int64, P],
    "smsut_sum_f32": [P, P, c_int64, c_float, P],
    "smsut_fill_f32": [P, c_int64, c_float, P],
    "smsut_fill_scaled_f32": [P, c_int64, P, c_float, P],
    "smsut_tanh_bwd": [P, P, P, c_int64, P],
    "smsut_lerp_rows_f32": [P, P, P, P, c_int, c_int64, P],
    "smsut_ce_rows_fwd": [P, P, P, c_int, c_int, c_float, P],
    "smsut_ce_rows_bwd": [P, P, P, c_float, P, c_int, c_int, P],
    "smsut_gp_fwd": [P, P, P, P, c_int, c_int64, c_float, P],
    "smsut_gp_bwd": [P, P, P, c_float, P, c_int, c_int64, P],
    "smsut_gather_rows": [P, P, P, c_int, c_int, c_int, c_int, P],
    "smsut_scatter_rows_add": [P, P, P, c_int, c_int, c_int, c_int, P],
    "smsut_l2norm_fwd": [P, P, P, c_int, c_int, P],
    "smsut_l2norm_bwd": [P, P, P, P, c_int, c_int, P],
    "smsut_patchnce_fwd": [P, P, P, P, c_int, c_int, c_int, c_float, c_float, P],
    "smsut_patchnce_bwd": [P, P, P, c_float, P, c_int, c_int, c_int, c_float, P],
    "smsut_sgd_step": [P, P, P, c_int64, P, c_float, c_float, c_float, P],
    "smsut_adam_step": [P, P, P, P, c_int64, P, c_float, c_float, c_float, c_float, P, c_float, P],
    "smsut_ema_update": [P, P, c_int64, P, P],
    "smsut_poly_lr_tick": [P, P, c_float, c_float, c_float, P],
    "smsut_pack_weights": [P, c_int, P],
    "smsut_unpack_wgrads": [P, c_int, P],
    "smsut_augment_batch": [P, P, P, P, P, P, c_int, c_int, c_int, P],
    "smsut_det_register": [P, C.c_size_t, P],
    "smsut_det_unregister": [P],
    "smsut_det_resolve": [P, c_int64, P],
}


class SmsutError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU / PyTorch fallback for the SMSUT hot path.")
    lib = C.CDLL(LIB_PATH)
    lib.smsut_last_error.restype = C.c_char_p
    lib.smsut_last_error.argtypes = []
    lib.smsut_abi_version.restype = c_int
    lib.smsut_launch_count.restype = c_int64
    lib.smsut_det_ranges.restype = c_int
    lib.smsut_det_ranges.argtypes = []
    lib.smsut_det_shadow.restype = c_void_p
    lib.smsut_det_shadow.argtypes = [c_void_p]
    for name, argtypes in _PROTOS.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = c_int
    return lib


lib = _load()


def check(rc, what=""):
    if rc != 0:
        msg = lib.smsut_last_error()
        raise SmsutError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")


def call(name, *args):
    check(getattr(lib, name)(*args), name)


def launch_count():
    return int(lib.smsut_launch_count())


def exported_names():
    return ["smsut_last_error", "smsut_abi_version", "smsut_launch_count", "smsut_det_ranges", "smsut_det_shadow"] + \
        list(_PROTOS)

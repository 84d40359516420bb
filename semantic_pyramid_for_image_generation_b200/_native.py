"""ctypes binding of libspyramid_b200.so (the C-ABI declared in include/spyramid_b200.h).

There is no fallback: if the shared library is missing or a call fails, a RuntimeError is raised.  Every wrapper
enqueues on torch's current CUDA stream, never synchronises, and is therefore CUDA-graph capturable.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libspyramid_b200.so")

c_void_p, c_int, c_float, c_ll, c_double = C.c_void_p, C.c_int, C.c_float, C.c_longlong, C.c_double


class ConvSrc(C.Structure):
    _fields_ = [("x", c_void_p), ("w", c_void_p), ("cin", c_int), ("ksize", c_int), ("w_mn_major", c_int),
                ("w_per_image", c_int), ("w_lo_off", c_ll)]


class ConvDesc(C.Structure):
    _fields_ = [("B", c_int), ("H", c_int), ("W", c_int), ("Cout", c_int), ("nsrc", c_int), ("src", ConvSrc * 9),
                ("bias", c_void_p), ("bias2", c_void_p), ("bias3", c_void_p), ("stencil_mask", c_void_p), ("stencil_w", c_void_p), ("dmask", c_void_p),
                ("dmask_slope", c_float), ("residual", c_void_p), ("y_raw", c_void_p), ("y_act", c_void_p),
                ("act", c_int), ("act_slope", c_float), ("y_f32", c_void_p), ("f32_store", c_int), ("splits", c_int),
                ("block_n", c_int), ("stages", c_int), ("residual_pooled", c_int), ("pool", c_int)]


class WgradDesc(C.Structure):
    _fields_ = [("B", c_int), ("H", c_int), ("W", c_int), ("Cin", c_int), ("Cout", c_int), ("ksize", c_int),
                ("x", c_void_p), ("dy", c_void_p), ("dw", c_void_p), ("cin_stride", c_int), ("splits", c_int),
                ("stages", c_int), ("per_image", c_int), ("dbg_lbo", c_int), ("dbg_sbo", c_int), ("scratch", c_void_p),
                ("scratch_floats", c_ll)]


class ReduceEntry(C.Structure):
    _fields_ = [("kind", c_int), ("partial", c_void_p), ("out0", c_void_p), ("out1", c_void_p), ("out2", c_void_p),
                ("nslices", c_int), ("nimg", c_int), ("taps", c_int), ("rows", c_int), ("cin_stride", c_int),
                ("Cout", c_int), ("lanes", c_int), ("nb", c_int), ("n", c_int), ("mode", c_int), ("C", c_int),
                ("sink_stride", c_int), ("sink_row", c_int), ("split", c_int)]


REDUCE_BATCH = 32  # SPYR_REDUCE_BATCH


class SnLayer(C.Structure):
    _fields_ = [("w", c_void_p), ("u", c_void_p), ("v", c_void_p), ("rows", c_int), ("cols", c_int), ("taps", c_int),
                ("cin", c_int), ("pack_cin", c_int), ("pack_mode", c_int), ("pack_off", c_ll), ("stencil_off", c_ll),
                ("gw_off", c_ll), ("gw_layout", c_int), ("grad_off", c_ll), ("index", c_int), ("tile0_wtu", c_int),
                ("tile0_wv", c_int), ("tile0_pack", c_int), ("tile0_bwd", c_int), ("tile0_tsum", c_int),
                ("scratch_off", c_ll), ("saved_off", c_ll), ("part_off", c_ll)]


class SnPlan(C.Structure):
    _fields_ = [("tiles_wtu", c_int), ("tiles_wv", c_int), ("tiles_pack", c_int), ("tiles_bwd", c_int),
                ("tiles_tsum", c_int), ("scratch_floats", c_ll), ("saved_floats", c_ll)]


ADAM_MAX_TENSORS = 48


class AdamChunk(C.Structure):
    _fields_ = [("count", c_int), ("p", c_void_p * ADAM_MAX_TENSORS), ("g", c_void_p * ADAM_MAX_TENSORS),
                ("m", c_void_p * ADAM_MAX_TENSORS), ("v", c_void_p * ADAM_MAX_TENSORS),
                ("n", c_ll * ADAM_MAX_TENSORS)]


P = c_void_p
_SIGNATURES = {
    "spyr_conv2d_fprop": [C.POINTER(ConvDesc), P],
    "spyr_conv2d_epilogue": [C.POINTER(ConvDesc), P, P],
    "spyr_conv2d_wgrad": [C.POINTER(WgradDesc), P],
    "spyr_conv2d_wgrad_deferred": [C.POINTER(WgradDesc), C.POINTER(ReduceEntry), P],
    "spyr_colsum_deferred": [P, c_ll, c_int, P, P, P, P, C.POINTER(ReduceEntry), P],
    "spyr_stencil_wgrad_deferred": [P, P, c_int, c_int, c_int, c_int, P, c_int, c_int, P, C.POINTER(ReduceEntry), P],
    "spyr_reduce_batched": [C.POINTER(ReduceEntry), c_int, P],
    "spyr_im2col3x3": [P, c_int, c_int, c_int, P, P, P, P],
    "spyr_col2im3x3": [P, c_int, c_int, c_int, P, P, c_int, P],
    "spyr_img_avgpool_pad8": [P, c_int, c_int, c_int, P, P],
    "spyr_img_avgpool_pad8_bwd": [P, c_int, c_int, c_int, P, c_int, P],
    "spyr_nchw_to_nhwc": [P, P, c_float, P, c_int, c_int, c_int, P],
    "spyr_nhwc_to_nchw": [P, P, c_float, P, c_int, c_int, c_int, P],
    "spyr_maskgate": [P, P, P, c_ll, c_int, P],
    "spyr_avgpool2_fwd": [P, P, P, P, c_float, c_int, c_int, c_int, c_int, P],
    "spyr_avgpool2_bwd": [P, P, c_int, c_int, c_int, c_int, P],
    "spyr_maxpool2_fwd": [P, P, c_int, c_int, c_int, c_int, P],
    "spyr_maxpool2_bwd": [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P],
    "spyr_adaptive_avgpool_fwd": [P, P, c_int, c_int, c_int, c_int, c_int, c_int, P],
    "spyr_adaptive_avgpool_bwd": [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P],
    "spyr_global_avgpool_lrelu_fwd": [P, c_float, P, c_int, c_int, c_int, P],
    "spyr_global_avgpool_lrelu_bwd": [P, P, c_float, P, c_int, c_int, c_int, P],
    "spyr_gamma_residual_fwd": [P, P, P, P, P, c_float, c_ll, P],
    "spyr_gamma_residual_bwd": [P, P, P, P, P, c_ll, P, P],
    "spyr_colsum": [P, c_ll, c_int, P, P, P, P, P],
    "spyr_stencil_wgrad": [P, P, c_int, c_int, c_int, c_int, P, c_int, c_int, P, P],
    "spyr_cast_f32_bf16": [P, P, c_ll, P],
    "spyr_vec_epilogue": [P, c_int, P, P, P, c_int, P, P, c_int, c_int, c_int, P],
    "spyr_conv1x1_tanh_fwd": [P, P, P, P, P, c_int, c_int, c_int, c_int, P],
    "spyr_conv1x1_tanh_bwd": [P, P, P, P, P, c_float, P, P, P, c_int, c_int, c_int, c_int, P, P],
    "spyr_bn_stats": [P, c_int, c_int, c_int, c_int, c_int, P, P],
    "spyr_up2_stats": [P, c_int, c_int, c_int, c_int, P, P, P],
    "spyr_bn_finalize": [P, c_double, c_int, c_float, c_float, P, P, P, P, c_int, P],
    "spyr_bn_act": [P, P, P, P, c_int, P, c_float, c_int, P, P, c_int, c_int, c_int, c_int, P],
    "spyr_bn_bwd_reduce": [P, P, P, P, P, c_int, P, c_float, c_int, P, P, c_int, c_int, c_int, c_int, P],
    "spyr_bn_bwd_finalize": [P, c_int, c_int, c_float, P, c_int, P, P, P, P, P],
    "spyr_bn_bwd_params": [P, c_int, c_int, c_int, P, P, P, P],
    "spyr_bn_bwd_apply": [P, P, P, P, c_int, P, P, P, P, c_int, c_int, c_int, c_int, c_int, P],  # B,H,W,C,x_up2
    "spyr_up2_bwd": [P, P, c_int, c_int, c_int, c_int, P],
    "spyr_argmax_rows": [P, c_int, c_int, c_int, P, P],
    "spyr_sn_plan": [C.POINTER(SnLayer), c_int, C.POINTER(SnPlan)],
    "spyr_sn_forward": [P, c_int, C.POINTER(SnPlan), c_int, c_float, P, P, P, P, P],
    "spyr_sn_backward": [P, c_int, C.POINTER(SnPlan), P, P, P, P, P],
    "spyr_linear_fwd": [P, P, c_float, P, P, P, P, c_float, P, c_int, c_int, c_int, P],
    "spyr_linear_bwd_x": [P, P, c_float, P, P, P, c_float, P, c_int, c_int, c_int, c_int, P],
    "spyr_linear_bwd_w": [P, P, c_float, P, P, c_float, P, P, c_int, c_int, c_int, P],
    "spyr_dhead_out_fwd": [P, P, P, P, P, P, c_int, c_int, P],
    "spyr_dhead_out_bwd": [P, P, P, P, P, P, P, P, c_int, c_int, P],
    "spyr_sagan_attention_fwd": [P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, P],
    "spyr_softmax_rows_fwd": [P, P, c_ll, c_int, P],
    "spyr_softmax_rows_bwd": [P, P, P, c_ll, c_int, P],
    "spyr_lsgan_fwd": [P, c_ll, c_float, P, P],
    "spyr_lsgan_bwd": [P, c_ll, c_float, P, P, P],
    "spyr_rec_level_fwd": [P, P, P, c_int, c_int, c_int, c_int, P, P, P],
    "spyr_rec_level_bwd": [P, P, P, c_int, c_int, c_int, c_int, P, P, P],
    "spyr_rec_vec_fwd": [P, P, P, c_int, c_int, P, P, P],
    "spyr_rec_vec_bwd": [P, P, P, c_int, c_int, P, P, P],
    "spyr_diversity_fwd": [P, c_ll, P, c_ll, P, P, P, P],
    "spyr_diversity_bwd": [P, c_ll, P, P, P, P],
    "spyr_add_inplace": [P, P, c_ll, P],
    "spyr_wgrad_to_oihw": [P, P, c_int, c_int, c_int, c_int, c_int, P],
    "spyr_dropout_fwd": [P, c_ll, c_float, C.c_ulonglong, C.c_ulonglong, P, P, P, P],
    "spyr_dropout_bwd": [P, P, c_ll, c_float, P, P],
    "spyr_weight_transpose_flip": [P, P, c_int, c_int, c_int, P],
    "spyr_expand_mask_level": [P, P, P, c_ll, c_int, c_int, c_int, c_int, P, P],
    "spyr_image_u8_minmax_normalize": [P, c_int, c_ll, P, P],
    "spyr_adam_tick": [P, P],
    "spyr_adam_step": [C.POINTER(AdamChunk), P, c_float, c_float, c_float, c_float, P],
}
EXPORTED_SYMBOLS = sorted(list(_SIGNATURES) + ["spyr_last_error", "spyr_version", "spyr_launch_count",
                                               "spyr_launch_count_reset", "spyr_set_precision", "spyr_get_precision",
                                               "spyr_conv2d_wgrad_scratch_floats", "spyr_last_conv_kernel"])
REDUCE_BLOCKS = 296  # SPYR_REDUCE_BLOCKS

_lib = None


def lib():
    """Loads the shared library on first use; raises if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                "libspyramid_b200.so is missing (%s). Build it with `make` or __graft_entry__.build(); this package "
                "has no CPU or PyTorch fallback for its CUDA kernels." % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        handle.spyr_last_error.restype = C.c_char_p
        handle.spyr_version.restype = c_int
        handle.spyr_launch_count.restype = c_ll
        handle.spyr_launch_count_reset.restype = None
        handle.spyr_last_conv_kernel.argtypes, handle.spyr_last_conv_kernel.restype = [], C.c_char_p
        handle.spyr_set_precision.argtypes, handle.spyr_set_precision.restype = [c_int], c_int
        handle.spyr_get_precision.argtypes, handle.spyr_get_precision.restype = [], c_int
        handle.spyr_conv2d_wgrad_scratch_floats.argtypes = [C.POINTER(WgradDesc)]
        handle.spyr_conv2d_wgrad_scratch_floats.restype = c_ll
        for name, argtypes in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.argtypes = argtypes
            fn.restype = c_int
        _lib = handle
    return _lib


def _check(rc, name):
    if rc != 0:
        raise RuntimeError("%s failed (rc=%d): %s" % (name, rc, lib().spyr_last_error().decode()))


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def call(name, *args):
    """Calls a C-ABI entry point with the current stream appended; raises RuntimeError on a non-zero status.
    Tensor arguments are passed as their data pointers (and stay alive for the duration of the enqueue)."""
    rc = getattr(lib(), name)(*[a.data_ptr() if isinstance(a, torch.Tensor) else a for a in args], stream_ptr())
    if rc != 0:
        _check(rc, name)


def call_nostream(name, *args):
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        _check(rc, name)


def launch_count():
    return int(lib().spyr_launch_count())


def launch_count_reset():
    lib().spyr_launch_count_reset()


def ptr(t):
    """data pointer of a tensor (None -> NULL); the tensor must be a CUDA tensor laid out as the kernel expects."""
    if t is None:
        return None
    return t.data_ptr()

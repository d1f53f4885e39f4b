"""ctypes binding of libgim_b200.so (the C ABI declared in include/gim_b200.h).

There is NO fallback: if the library is missing or a call fails, a RuntimeError is raised.  PyTorch only supplies
device memory (`tensor.data_ptr()`) and the current stream.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgim_b200.so")

F32, BF16 = 0, 1
ALGO_AUTO, ALGO_SIMT, ALGO_TCGEN05 = 0, 1, 2

_P, _I, _L, _F = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float
_CODES = {"p": _P, "i": _I, "l": _L, "f": _F}

# name -> argument codes (must match include/gim_b200.h; tests/test_cabi_symbols.py checks the symbol list)
PROTOTYPES = {
    "gim_conv2d_fwd": "ppppiiiiiiiiip",
    "gim_conv2d_fwd_fused": "ppppppiiiiiiiifp",
    "gim_pool2_multi": "pppppiiiiffp",
    "gim_first_block_fwd": "pppppppiiiiiifp",
    "gim_first_block_wgrad": "ppppppliiiiiifp",
    "gim_unpool2_cast": "ppiiiifp",
    "gim_conv2d_wgrad": "pppiiiiiiiip",
    "gim_weight_cast": "ppiiiip",
    "gim_weight_flip": "ppiiiip",
    "gim_colsum": "pplii" + "p",
    "gim_colsum_acc": "pplii" + "p",
    "gim_cast_colsum": "pppliip",
    "gim_sn_forward": "pppifpppppiiip",
    "gim_sn_backward": "pppppppiiip",
    "gim_sn_forward_multi": "piifp",
    "gim_sn_backward_multi": "pip",
    "gim_lrelu_fwd": "pplfip",
    "gim_lrelu_bwd": "ppplfip",
    "gim_tanh_fwd": "pplip",
    "gim_tanh_bwd": "ppplip",
    "gim_axpby": "ppplffip",
    "gim_scale_dev": "ppplip",
    "gim_dot": "ppplip",
    "gim_pool2_sum": "pppiiiifip",
    "gim_unpool2_bcast": "ppiiiifip",
    "gim_nchw_to_nhwc": "ppiiiiip",
    "gim_nhwc_to_nchw": "ppiiiiip",
    "gim_copy_cols": "piipiiliip",
    "gim_cast": "pipilp",
    "gim_operand_prepare": "pipiiiiiifp",
    "gim_lrelu_bwd_ref": "ppiplfp",
    "gim_im2col": "ppiiiiiiiip",
    "gim_col2im": "pppiiiiiip",
    "gim_norm_stats": "pppiiiip",
    "gim_affine_act_fwd": "pppppiiifip",
    "gim_norm_bwd_reduce": "ppppppppiiifip",
    "gim_norm_bwd_apply": "ppppppppppiiifip",
    "gim_norm_act_operand": "pppppiiiifip",
    "gim_norm_coeffs": "ippppppiiifp",
    "gim_norm_bwd_coeffs": "ippppppppp" + "iiifp",
    "gim_gemm_strided": "pilllpilllpilliiiiffp",
    "gim_gemm_strided_bf16": "pilllpilllpilliiiiffp",
    "gim_bias_act_fwd": "ppplifp",
    "gim_softmax_rows_fwd": "pplip",
    "gim_softmax_rows_bwd": "ppplip",
    "gim_softmax_rows_bwd_bwd": "pppplip",
    "gim_attention_fwd": "pppiippppiiip",
    "gim_attention_bwd": "ppppiippppppiiip",
    "gim_set_stats_fwd": "pppiiiiffp",
    "gim_set_stats_bwd": "ppippiiiffp",
    "gim_set_std_bwd_bwd": "ppippp" + "iiifp",
    "gim_set_center_add": "pppiiiip",
    "gim_img_att_blend_fwd": "pppppppplip",
    "gim_img_att_blend_bwd": "pppppppppppppplip",
    "gim_affine_rows": "pppiiiffp",
    "gim_gmax_fwd": "pppiiifip",
    "gim_gather_idx": "pppiiiip",
    "gim_scatter_idx": "pppiiiip",
    "gim_bce_logits_fwd": "pfplp",
    "gim_bce_logits_bwd": "ppfplp",
    "gim_rows_sqsum": "ppilip",
    "gim_rows_scale": "pppilfip",
    "gim_adam_multi": "pilppfff" + "fp",
    "gim_zero_grads_multi": "pilp",
}
OTHER_SYMBOLS = ("gim_version", "gim_last_error", "gim_conv2d_tc_supported", "gim_conv2d_wgrad_tc_supported", "gim_conv2d_fwd_plan", "gim_conv2d_wgrad_plan", "gim_launch_count",
                 "gim_set_deterministic")



class SnLayer(ctypes.Structure):
    """gim_sn_layer of include/gim_b200.h."""
    _fields_ = [("w", _P), ("u", _P), ("v", _P), ("w_sn", _P), ("w_op", _P), ("w_flip", _P), ("aux", _P), ("scratch", _P),
                ("cout", _I), ("cin", _I), ("ksize", _I), ("reserved", _I)]


class SnBwdLayer(ctypes.Structure):
    """gim_sn_bwd_layer of include/gim_b200.h."""
    _fields_ = [("g", _P), ("w", _P), ("u", _P), ("v", _P), ("sigma", _P), ("grad", _P), ("scratch", _P),
                ("cout", _I), ("cin", _I), ("ksize", _I), ("accumulate", _I)]


_lib = None


def lib():
    """Load libgim_b200.so (once).  Raises if it was not built: the product path has no CPU fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libgim_b200.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C optimalstrategiesagainstgenerativeattacks_b200/csrc`). There is no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, codes in PROTOTYPES.items():
            fn = getattr(L, name)
            fn.argtypes = [_CODES[c] for c in codes]
            fn.restype = _I
        L.gim_version.restype = _I
        L.gim_last_error.restype = ctypes.c_char_p
        L.gim_conv2d_tc_supported.argtypes = [_I] * 7
        L.gim_conv2d_tc_supported.restype = _I
        L.gim_conv2d_wgrad_tc_supported.argtypes = [_I] * 7
        L.gim_conv2d_wgrad_tc_supported.restype = _I
        L.gim_conv2d_fwd_plan.argtypes = [_I] * 8 + [_P]
        L.gim_conv2d_fwd_plan.restype = _I
        L.gim_conv2d_wgrad_plan.argtypes = [_I] * 6 + [_P]
        L.gim_conv2d_wgrad_plan.restype = _I
        L.gim_launch_count.argtypes = [_I]
        L.gim_launch_count.restype = _L
        L.gim_set_deterministic.argtypes = [_I]
        L.gim_set_deterministic.restype = _I
        _lib = L
    return _lib


def dtype_code(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError("libgim_b200 handles float32 / bfloat16 tensors only, got %s" % t.dtype)


def torch_dtype(code):
    return torch.float32 if code == F32 else torch.bfloat16


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("libgim_b200: tensor is on %s; the GIM hot path runs on CUDA only (no CPU fallback)" % t.device)
    if not t.is_contiguous():
        raise RuntimeError("libgim_b200: tensor must be contiguous")
    if t.device.index != _current_device():
        raise RuntimeError("libgim_b200: tensor lives on %s but the current CUDA device is %d -- kernels launch on the current device's "
                           "stream; call torch.cuda.set_device() (utils.get_device does) or run one process per GPU" % (t.device, _current_device()))
    return t.data_ptr()


_current_device = torch.cuda.current_device


def stream():
    return torch.cuda.current_stream().cuda_stream


def call(name, *args):
    L = lib()
    rc = getattr(L, name)(*args, stream())
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (name, rc, L.gim_last_error().decode()))


def launch_count(reset=False):
    return int(lib().gim_launch_count(1 if reset else 0))


def set_deterministic(flag):
    """Parity mode (gim_set_deterministic): no multi-CTA fp32 atomics into one output.  -> previous setting."""
    return bool(lib().gim_set_deterministic(1 if flag else 0))


def conv_tc_supported(n, h, w, cin, cout, k, dtype):
    return bool(lib().gim_conv2d_tc_supported(n, h, w, cin, cout, k, dtype))


PLAN_FIELDS = ("persistent", "halo", "block_n", "m_sub", "pair", "stages", "a_stages", "b_stages", "k_chains", "bw", "bh", "bn", "grid_x", "grid_y",
               "threads", "smem_bytes", "tmem_cols", "block_k", "pixel_tiles", "halo_bytes")


def conv_fwd_plan(n, h, w, cin, cout, k, out_dtype=F32, epilogue=0):
    """The launch configuration the tensor-core forward / input-gradient path would use for this shape (no device needed)."""
    buf = (ctypes.c_int * 20)()
    rc = lib().gim_conv2d_fwd_plan(n, h, w, cin, cout, k, out_dtype, epilogue, ctypes.cast(buf, ctypes.c_void_p))
    if rc != 0:
        raise RuntimeError(lib().gim_last_error().decode())
    return dict(zip(PLAN_FIELDS, list(buf)))


WGRAD_PLAN_FIELDS = ("kernel", "block_n", "stages", "grid_x", "grid_y", "threads", "smem_bytes", "tmem_cols", "pixel_tiles", "tiles_per_split",
                     "taps_per_cta", "bw", "bh", "bn")


def conv_wgrad_plan(n, h, w, cin, cout, k):
    """The launch configuration of the tensor-core weight gradient for this shape (no device needed)."""
    buf = (ctypes.c_int * 16)()
    rc = lib().gim_conv2d_wgrad_plan(n, h, w, cin, cout, k, ctypes.cast(buf, ctypes.c_void_p))
    if rc != 0:
        raise RuntimeError(lib().gim_last_error().decode())
    return dict(zip(WGRAD_PLAN_FIELDS, list(buf)))


def wgrad_tc_supported(n, h, w, cin, cout, k, dtype):
    return bool(lib().gim_conv2d_wgrad_tc_supported(n, h, w, cin, cout, k, dtype))

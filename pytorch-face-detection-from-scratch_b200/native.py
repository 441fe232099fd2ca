"""ctypes binding of libfd_b200.so (the C ABI declared in include/fd_b200.h).

There is deliberately NO fallback: if the library is missing or a tensor is not a contiguous
CUDA tensor the call raises.  The product path never imports ``oracle/``.
"""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfd_b200.so")

_c = ctypes
_P, _I, _F, _D = _c.c_void_p, _c.c_int, _c.c_float, _c.c_double

# name -> argtypes, in the order of include/fd_b200.h (restype is int unless noted)
SIGNATURES = {
    "fd_version": [],
    "fd_launch_count": [],
    "fd_error_string": [_I],
    "fd_conv3x3": [_P, _P, _I, _I, _I, _I, _P, _F, _P, _P, _P, _P, _P, _P, _P, _I, _P],
    "fd_conv3x3_pool": [_P, _P, _I, _I, _I, _I, _P, _F, _P, _P, _P, _P, _P, _I, _P],
    "fd_conv3x3_wide": [_P, _I, _P, _I, _I, _I, _P, _F, _P, _P, _P, _P, _P, _P, _P, _I, _P],
    "fd_conv3x3_wide_shared_tile": [_I, _I, _I, _I],
    "fd_conv3x3_wide_chain_ok": [_I, _I, _I],
    "fd_conv3x3_wide_chain": [_P, _I, _P, _I, _I, _I, _I, _F, _P, _I, _P],
    "fd_pack_conv3x3_wide": [_P, _I, _I, _I, _P, _P, _P],
    "fd_pack_conv1x1_wide": [_P, _I, _I, _I, _P, _P, _P],
    "fd_conv3x3_wgrad_wide": [_P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _c.c_long, _P, _P, _c.c_long, _I, _P],
    "fd_conv3x3_wgrad": [_P, _P, _I, _I, _I, _I, _P, _P, _I, _P],
    "fd_conv3x3_wgrad_multi": [_P, _P, _I, _I, _I, _I, _I, _P, _c.c_long, _P, _c.c_long, _I, _P],
    "fd_resblock_chain_shape_ok": [_I, _I, _I],
    "fd_resblock_chain_fwd": [_P, _P, _P, _I, _I, _I, _I, _I, _F, _P],
    "fd_resblock_chain_bwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P],
    "fd_pack_conv3x3": [_P, _I, _I, _P, _P, _P],
    "fd_unpack_wgrad3x3": [_P, _I, _I, _P, _P],
    "fd_unpack_wgrad3x3_planes": [_P, _I, _I, _P, _P],
    "fd_adam_flat": [_P, _P, _P, _P, _c.c_long, _F, _F, _F, _F, _F, _I, _P, _P],
    "fd_comm_window_bytes": [_c.c_long, _I],
    "fd_comm_error_offset": [],
    "fd_comm_status": [_P, _P],
    "fd_comm_alloc": [_c.c_long, _P],
    "fd_comm_free": [_P],
    "fd_comm_export": [_P, _P],
    "fd_comm_import": [_P, _P],
    "fd_comm_release": [_P],
    "fd_allreduce_sum_f32": [_P, _I, _I, _P, _c.c_long, _P],
    "fd_allreduce_sum_f32_blocks": [_P, _I, _I, _P, _c.c_long, _I, _P],
    "fd_dwconv3x3_lrelu": [_P, _P, _I, _I, _I, _I, _F, _P, _P],
    "fd_act_mask": [_P, _I, _I, _I, _F, _P, _P, _P, _P, _P],
    "fd_grad_mask": [_P, _I, _I, _I, _F, _P, _P, _P, _P],
    "fd_dropout_scale": [_P, _c.c_long, _c.c_long, _F, _F, _P, _P],
    "fd_sepblock_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _F, _I, _P, _P],
    "fd_sep_pack": [_P, _c.c_long, _P, _P, _I, _P, _P],
    "fd_stem_cache_elems": [_I, _I, _I, _I, _I, _I, _I, _I],
    "fd_stem_fwd": [_P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P],
    "fd_stem_wgrad": [_P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P],
    "fd_stem_fwd_cached": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P],
    "fd_stem_wgrad_pair": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P],
    "fd_head_pack": [_P, _I, _I, _P, _P],
    "fd_head_fwd": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P],
    "fd_head_bwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _F, _P, _P, _P, _P],
    "fd_maxpool2x2_fwd": [_P, _I, _I, _I, _I, _P, _P, _P],
    "fd_maxpool2x2_bwd": [_P, _P, _I, _I, _I, _I, _P, _P, _P, _F, _P, _P, _P],
    "fd_yolo_loss": [_P, _P, _I, _I, _I, _P, _P, _P, _P],
    "fd_decode_nms": [_P, _I, _I, _I, _F, _D, _I, _I, _I, _P, _P, _P, _P],
    "fd_box_metrics": [_P, _P, _P, _P, _I, _I, _F, _P, _P],
    "fd_grid_encode": [_P, _P, _I, _I, _I, _I, _P, _P],
    "fd_ssd_grid_encode": [_P, _P, _I, _P, _I, _I, _I, _P, _P],
    "fd_ssd_decode_nms": [_P, _I, _P, _I, _F, _D, _I, _I, _I, _P, _P, _P],
    "fd_ssd_loss": [_P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P],
    "fd_pw_packed_elems": [_I, _I],
    "fd_pw_padded_n": [_I],
    "fd_pw_pack": [_P, _P, _I, _I, _P, _P],
    "fd_pw_conv": [_P, _P, _P, _c.c_long, _I, _I, _I, _P, _P, _P],
    "fd_mbv3_stem": [_P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P],
    "fd_dw_pack": [_P, _P, _I, _I, _P, _P],
    "fd_dwconv": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P],
    "fd_se_gate": [_P, _I, _I, _I, _P, _P, _P, _P, _I, _I, _P, _P],
    "fd_dwconv_se_blocks": [_I, _I, _I],
    "fd_scale_channels": [_P, _P, _I, _I, _I, _P],
    "fd_head3x3_fwd": [_P, _P, _P, _I, _I, _I, _I, _P, _P],
    "fd_resize_bilinear": [_P, _I, _c.c_long, _I, _I, _I, _I, _P, _P],
    "fd_index_copy_f32": [_P, _P, _P, _c.c_long, _I, _P],
    "fd_lrelu_bwd": [_P, _P, _c.c_long, _F, _P, _P],
    "fd_dwconv3x3_wgrad": [_P, _P, _I, _I, _I, _I, _P, _P],
    "fd_ssd_head_fwd": [_P, _I, _P, _P, _I, _I, _I, _P, _P, _I, _I, _P, _P],
    "fd_ssd_head_bwd": [_P, _P, _I, _P, _I, _I, _I, _P, _I, _I, _P, _P, _P, _P, _P],
}



class ChainFwdBlock(_c.Structure):
    """fd_chain_fwd_block of include/fd_b200.h"""
    _fields_ = [("bias1", _P), ("bias2", _P), ("chan_scale", _P), ("a", _P), ("mask_a", _P), ("mask_b", _P),
                ("out", _P)]


class ChainBwdBlock(_c.Structure):
    """fd_chain_bwd_block of include/fd_b200.h"""
    _fields_ = [("mask_a", _P), ("gp1", _P), ("g_in", _P), ("mask_b_prev", _P), ("chan_scale_prev", _P),
                ("gp2_prev", _P)]


class WideChainLayer(_c.Structure):
    """fd_wide_chain_layer of include/fd_b200.h"""
    _fields_ = [("in_index", _I), ("w_index", _I), ("flags", _I), ("reserved", _I), ("bias", _P), ("residual", _P * 2),
                ("out", _P * 2), ("out2", _P * 2), ("chan_scale", _P * 2), ("chan_scale2", _P * 2), ("mask_in", _P * 2),
                ("mask_out", _P * 2)]


_lib = None


class NativeError(RuntimeError):
    pass


def lib():
    """Load libfd_b200.so (once).  Raises NativeError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or python pytorch-face-detection-from-scratch_b200/csrc/build.py). There is no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError here == header / library out of sync
            fn.argtypes = args
            fn.restype = _I
        L.fd_launch_count.restype = _c.c_longlong
        L.fd_stem_cache_elems.restype = _c.c_long
        L.fd_comm_window_bytes.restype = _c.c_long
        L.fd_pw_packed_elems.restype = _c.c_long
        L.fd_error_string.restype = _c.c_char_p
        _lib = L
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().fd_error_string(rc).decode()
        raise NativeError(f"{what} failed: {msg} (code {rc})")


def launch_count() -> int:
    return int(lib().fd_launch_count())


def dptr(t, dtype=None):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise NativeError("fd_b200 kernels take CUDA tensors only (no CPU fallback); got a CPU tensor")
    if not t.is_contiguous():
        raise NativeError("fd_b200 kernels take contiguous tensors")
    if dtype is not None and t.dtype != dtype:
        raise NativeError(f"expected dtype {dtype}, got {t.dtype}")
    if t.device.index != torch.cuda.current_device():
        # kernels are enqueued on the CURRENT device's stream: a tensor of another device would fault asynchronously
        raise NativeError(f"tensor lives on cuda:{t.device.index} but the current device is "
                          f"cuda:{torch.cuda.current_device()}; wrap the call in torch.cuda.device(tensor.device)")
    return t.data_ptr()


def cur_stream():
    return torch.cuda.current_stream().cuda_stream

"""Typed Python wrappers over the C ABI -- one function per entry point of include/fd_b200.h.
Tensors are caller-allocated; nothing here allocates or synchronises."""
from __future__ import annotations

import torch

import ctypes

from .native import ChainBwdBlock, ChainFwdBlock, WideChainLayer, check, cur_stream, dptr, lib

BF16, F32, I32, U8 = torch.bfloat16, torch.float32, torch.int32, torch.uint8

EPI_LRELU = 1
CONV_1X1 = 2          # fd_conv3x3: centre tap only (FD_CONV_1X1)
CONV_ONE_TAP = 4      # fd_conv3x3: the one-tap-per-MMA kernel (FD_CONV_ONE_TAP); the chain kernels are bit-identical to it


def conv3x3(x, w_packed, *, bias=None, slope=0.2, lrelu=False, chan_scale=None, residual=None, mask_out=None,
            out=None, mask_in=None, chan_scale2=None, out2=None, flags=0):
    """mask_out / mask_in: int32 [B,H,W,C//32] sign-bit masks (see include/fd_b200.h)."""
    B, H, W, C = x.shape
    fl = flags | (EPI_LRELU if lrelu else 0)
    check(lib().fd_conv3x3(dptr(x, BF16), dptr(w_packed, BF16), B, H, W, C, dptr(bias, F32), slope,
                           dptr(chan_scale, F32), dptr(residual, BF16), dptr(mask_out, I32), dptr(out, BF16),
                           dptr(mask_in, I32), dptr(chan_scale2, F32), dptr(out2, BF16), fl, cur_stream()),
          "fd_conv3x3")


def conv3x3_pool(x, w_packed, pooled, *, bias=None, slope=0.2, lrelu=True, chan_scale=None, residual=None, mask_out=None,
                 argmax=None, flags=0):
    """conv3x3 (+ bias, LeakyReLU, Dropout2d multiplier, sign-bit mask, skip add) with the MaxPool2d(2) that follows
    fused: pooled [B,H/2,W/2,C]; argmax int16 [B,H/2,W/2,C/8] (2-bit window positions) or None.  Even H, W only."""
    B, H, W, C = x.shape
    fl = flags | (EPI_LRELU if lrelu else 0)
    check(lib().fd_conv3x3_pool(dptr(x, BF16), dptr(w_packed, BF16), B, H, W, C, dptr(bias, F32), slope,
                                dptr(chan_scale, F32), dptr(residual, BF16), dptr(mask_out, I32), dptr(pooled, BF16),
                                dptr(argmax, torch.int16), fl, cur_stream()), "fd_conv3x3_pool")


def _plane_ptrs(tensors, dtype):
    """ctypes array of per-plane device pointers (None -> NULL array)."""
    if tensors is None:
        return None
    return (ctypes.c_void_p * len(tensors))(*[dptr(t, dtype) for t in tensors])


def conv3x3_wide(x_planes, w_packed, *, bias=None, slope=0.2, lrelu=False, chan_scale=None, residual=None, mask_out=None,
                 out=None, mask_in=None, chan_scale2=None, out2=None, flags=0):
    """fd_conv3x3_wide: 64*len(x_planes) -> 128 channels on [B,H,W,64] planes; every per-plane argument is a list of two
    tensors (x_planes: gin tensors).  w_packed: [gin,9,128,64] bf16 (ops.pack_conv3x3_wide, one group of 128 couts)."""
    B, H, W, C = x_planes[0].shape
    assert C == 64
    fl = flags | (EPI_LRELU if lrelu else 0)
    keep = [_plane_ptrs(x_planes, BF16), _plane_ptrs(chan_scale, F32), _plane_ptrs(residual, BF16), _plane_ptrs(mask_out, I32),
            _plane_ptrs(out, BF16), _plane_ptrs(mask_in, I32), _plane_ptrs(chan_scale2, F32), _plane_ptrs(out2, BF16)]
    cast = [ctypes.cast(a, ctypes.c_void_p) if a is not None else None for a in keep]
    check(lib().fd_conv3x3_wide(cast[0], len(x_planes), dptr(w_packed, BF16), B, H, W, dptr(bias, F32), slope, cast[1],
                                cast[2], cast[3], cast[4], cast[5], cast[6], cast[7], fl, cur_stream()), "fd_conv3x3_wide")


def conv3x3_wide_shared_tile(B, H, W, flags=0):
    """True when fd_conv3x3_wide runs this map in shared-tile mode, i.e. may write out AND out2 in one launch."""
    return bool(lib().fd_conv3x3_wide_shared_tile(int(B), int(H), int(W), int(flags)))


def conv3x3_wide_chain_ok(B, H, W):
    """True when a run of wide convolutions on this map can go out as ONE launch (a CTA pair holds a whole image)."""
    return bool(lib().fd_conv3x3_wide_chain_ok(int(B), int(H), int(W)))


def wide_chain_layer(in_index, w_index, *, bias=None, lrelu=False, chan_scale=None, residual=None, mask_out=None, out=None,
                     mask_in=None, chan_scale2=None, out2=None):
    """One fd_wide_chain_layer; per-plane arguments are lists of two tensors as in conv3x3_wide."""
    L = WideChainLayer()
    L.in_index, L.w_index, L.flags, L.reserved = int(in_index), int(w_index), (EPI_LRELU if lrelu else 0), 0
    L.bias = dptr(bias, F32)
    for name, planes, dt in (("residual", residual, BF16), ("out", out, BF16), ("out2", out2, BF16), ("chan_scale", chan_scale, F32),
                             ("chan_scale2", chan_scale2, F32), ("mask_in", mask_in, I32), ("mask_out", mask_out, I32)):
        arr = getattr(L, name)
        for g in range(2):
            arr[g] = dptr(planes[g], dt) if planes is not None else None
    return L


def conv3x3_wide_chain(x_stack, w_packed, layers, slope=0.2):
    """fd_conv3x3_wide_chain.  x_stack: two stacked plane buffers [n_stack,B,H,W,64] bf16; w_packed: [n_w_layers,1,2,9,128,64]
    (or [n_w_layers,2,9,128,64]) bf16; layers: wide_chain_layer(...) records, executed in order by one launch."""
    n_stack, B, H, W, C = x_stack[0].shape
    assert C == 64 and x_stack[1].shape == x_stack[0].shape and w_packed.numel() % (2 * 9 * 128 * 64) == 0
    arr = (WideChainLayer * len(layers))(*layers)
    ptrs = (ctypes.c_void_p * 2)(dptr(x_stack[0], BF16), dptr(x_stack[1], BF16))
    check(lib().fd_conv3x3_wide_chain(ctypes.cast(ptrs, ctypes.c_void_p), n_stack, dptr(w_packed, BF16),
                                      w_packed.numel() // (2 * 9 * 128 * 64), B, H, W, slope, ctypes.cast(arr, ctypes.c_void_p),
                                      len(layers), cur_stream()), "fd_conv3x3_wide_chain")


def pack_conv3x3_wide(w, w_fwd, w_dgrad):
    """w: [L,Cout,Cin,k,k] (or [Cout,Cin,k,k]) fp32, k = 3 or 1 -> w_fwd [L,Cout/128,Cin/64,9,128,64], w_dgrad
    [L,Cin/128,Cout/64,9,128,64] (k = 1: only the centre tap is written -- for CONV_1X1)."""
    if w.dim() == 5:
        n, Cout, Cin, k = w.shape[0], w.shape[1], w.shape[2], w.shape[3]
    else:
        n, Cout, Cin, k = 1, w.shape[0], w.shape[1], w.shape[2]
    fn = lib().fd_pack_conv3x3_wide if k == 3 else lib().fd_pack_conv1x1_wide
    check(fn(dptr(w, F32), n, Cout, Cin, dptr(w_fwd, BF16), dptr(w_dgrad, BF16), cur_stream()), "fd_pack_conv_wide")


def conv3x3_wgrad_wide(x0, x1, g0, g1, dw_packed, sub_off, dw_stride=0, dbias0=None, dbias1=None, dbias_stride=0, flags=0):
    """fd_conv3x3_wgrad_wide: the 128 x 128 channel block (x planes x0, x1) x (gradient planes g0, g1) of the weight
    gradient; tensors [B,H,W,64] or stacked [nprob,B,H,W,64].  sub_off: 4 element offsets into dw_packed of the packed
    [9,64,64] sub-blocks (x0,g0), (x0,g1), (x1,g0), (x1,g1)."""
    if x0.dim() == 5:
        nprob, B, H, W, C = x0.shape
    else:
        nprob = 1
        B, H, W, C = x0.shape
    assert C == 64 and len(sub_off) == 4
    offs = (ctypes.c_long * 4)(*[int(o) for o in sub_off])
    check(lib().fd_conv3x3_wgrad_wide(dptr(x0, BF16), dptr(x1, BF16), dptr(g0, BF16), dptr(g1, BF16), nprob, B, H, W,
                                      dptr(dw_packed, F32), ctypes.cast(offs, ctypes.c_void_p), int(dw_stride),
                                      dptr(dbias0, F32), dptr(dbias1, F32), int(dbias_stride), flags, cur_stream()),
          "fd_conv3x3_wgrad_wide")


def conv3x3_wgrad(x, g, dw_packed, dbias, flags=0):
    B, H, W, C = x.shape
    check(lib().fd_conv3x3_wgrad(dptr(x, BF16), dptr(g, BF16), B, H, W, C, dptr(dw_packed, F32), dptr(dbias, F32),
                                 flags, cur_stream()), "fd_conv3x3_wgrad")


def conv3x3_wgrad_multi(x, g, dw_packed, dw_stride, dbias, dbias_stride, flags=0):
    """x, g: [nprob,B,H,W,C] stacked problems; problem q accumulates into dw_packed[q*dw_stride:], dbias[q*dbias_stride:]."""
    nprob, B, H, W, C = x.shape
    check(lib().fd_conv3x3_wgrad_multi(dptr(x, BF16), dptr(g, BF16), nprob, B, H, W, C, dptr(dw_packed, F32),
                                       int(dw_stride), dptr(dbias, F32), int(dbias_stride), flags, cur_stream()),
          "fd_conv3x3_wgrad_multi")


def unpack_wgrad3x3_planes(dw_packed, G, dw):
    """dw_packed: [L*G*G, 9, 64, 64] fp32 sub-blocks (g-major, then h) -> dw: [L, 64G, 64G, 3, 3] fp32."""
    check(lib().fd_unpack_wgrad3x3_planes(dptr(dw_packed, F32), dw.shape[0], int(G), dptr(dw, F32), cur_stream()),
          "fd_unpack_wgrad3x3_planes")


def resblock_chain_ok(H, W, C):
    return bool(lib().fd_resblock_chain_shape_ok(int(H), int(W), int(C)))


def resblock_chain_fwd(x, w_fwd, blocks, slope=0.2):
    """blocks: list of dicts with keys bias1, bias2 and optional chan_scale, a, mask_a, mask_b, out (tensors).
    w_fwd: [2*len(blocks), 9, C, C] bf16 forward-packed weights of the run."""
    B, H, W, C = x.shape
    arr = (ChainFwdBlock * len(blocks))()
    for i, d in enumerate(blocks):
        arr[i] = ChainFwdBlock(dptr(d["bias1"], F32), dptr(d["bias2"], F32), dptr(d.get("chan_scale"), F32),
                               dptr(d.get("a"), BF16), dptr(d.get("mask_a"), I32), dptr(d.get("mask_b"), I32),
                               dptr(d.get("out"), BF16))
    check(lib().fd_resblock_chain_fwd(dptr(x, BF16), dptr(w_fwd, BF16), ctypes.cast(arr, ctypes.c_void_p),
                                      len(blocks), B, H, W, C, slope, cur_stream()), "fd_resblock_chain_fwd")


def resblock_chain_bwd(g_out, gp2_last, w_dgrad, blocks, slope=0.2):
    """blocks (FORWARD order): dicts with mask_a and optional gp1, g_in, mask_b_prev, chan_scale_prev, gp2_prev."""
    B, H, W, C = g_out.shape
    arr = (ChainBwdBlock * len(blocks))()
    for i, d in enumerate(blocks):
        arr[i] = ChainBwdBlock(dptr(d["mask_a"], I32), dptr(d.get("gp1"), BF16), dptr(d.get("g_in"), BF16),
                               dptr(d.get("mask_b_prev"), I32), dptr(d.get("chan_scale_prev"), F32),
                               dptr(d.get("gp2_prev"), BF16))
    check(lib().fd_resblock_chain_bwd(dptr(g_out, BF16), dptr(gp2_last, BF16), dptr(w_dgrad, BF16),
                                      ctypes.cast(arr, ctypes.c_void_p), len(blocks), B, H, W, C, slope,
                                      cur_stream()), "fd_resblock_chain_bwd")


def pack_conv3x3(w, w_fwd, w_dgrad):
    n, C = (w.shape[0], w.shape[1]) if w.dim() == 5 else (1, w.shape[0])
    check(lib().fd_pack_conv3x3(dptr(w, F32), n, C, dptr(w_fwd, BF16), dptr(w_dgrad, BF16), cur_stream()),
          "fd_pack_conv3x3")


def unpack_wgrad3x3(dw_packed, dw):
    n, C = (dw.shape[0], dw.shape[1]) if dw.dim() == 5 else (1, dw.shape[0])
    check(lib().fd_unpack_wgrad3x3(dptr(dw_packed, F32), n, C, dptr(dw, F32), cur_stream()), "fd_unpack_wgrad3x3")


def adam_flat(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, state=None):
    """state: int32[4] device tensor (step count, ticket, lr bits, -) for graph-replayable launches, else None."""
    check(lib().fd_adam_flat(dptr(p, F32), dptr(g, F32), dptr(m, F32), dptr(v, F32), p.numel(), float(lr), float(beta1),
                             float(beta2), float(eps), float(weight_decay), int(step), dptr(state, I32), cur_stream()),
          "fd_adam_flat")


def dwconv3x3_lrelu(x, w_dw, slope, out):
    """64-channel plane: out = lrelu(depthwise3x3(x)); w_dw [9,64] fp32 (fd_sep_pack layout)."""
    B, H, W, C = x.shape
    check(lib().fd_dwconv3x3_lrelu(dptr(x, BF16), dptr(w_dw, F32), B, H, W, C, float(slope), dptr(out, BF16), cur_stream()),
          "fd_dwconv3x3_lrelu")


def act_mask(x, slope, chan_scale, residual, mask_out, out):
    """64-channel plane: out = lrelu(x) * chan_scale + residual, mask_out = sign bits of x (nullable operands)."""
    B, H, W, C = x.shape
    check(lib().fd_act_mask(dptr(x, BF16), B, H * W, C, float(slope), dptr(chan_scale, F32), dptr(residual, BF16),
                            dptr(mask_out, I32), dptr(out, BF16), cur_stream()), "fd_act_mask")


def grad_mask(g, slope, mask_bits, chan_scale, out):
    """64-channel plane: out = g * (mask bit ? 1 : slope) * chan_scale."""
    B, H, W, C = g.shape
    check(lib().fd_grad_mask(dptr(g, BF16), B, H * W, C, float(slope), dptr(mask_bits, I32), dptr(chan_scale, F32),
                             dptr(out, BF16), cur_stream()), "fd_grad_mask")


def dropout_scale(r, n_block, keep_block, keep_head, out):
    check(lib().fd_dropout_scale(dptr(r, F32), r.numel(), int(n_block), float(keep_block), float(keep_head),
                                 dptr(out, F32), cur_stream()), "fd_dropout_scale")


def stem_cache_elems(B, Cin, Hin, Win, C, K, stride, pad):
    return int(lib().fd_stem_cache_elems(B, Cin, Hin, Win, C, K, stride, pad))


def sepblock_fwd(x, w_pw1, w_dw, w_pw2, slope, out, pool=False):
    """x: [B,H,W,64] bf16; out: [B,H,W,64] (or [B,H/2,W/2,64] with the fused MaxPool2d(2)); w_pw1 / w_pw2: [64,64]
    bf16; w_dw: [9,64] fp32 (fd_sep_pack layouts)."""
    B, H, W, C = x.shape
    assert tuple(out.shape) == ((B, H // 2, W // 2, C) if pool else (B, H, W, C))
    check(lib().fd_sepblock_fwd(dptr(x, BF16), dptr(w_pw1, BF16), dptr(w_dw, F32), dptr(w_pw2, BF16), B, H, W, C,
                                float(slope), int(bool(pool)), dptr(out, BF16), cur_stream()), "fd_sepblock_fwd")


def sep_pack(pw, pw_out, dw, dw_out):
    """pw: [L,64,64] fp32 -> pw_out bf16 (same shape); dw: [L,64,3,3] fp32 -> dw_out [L,9,64] fp32."""
    check(lib().fd_sep_pack(dptr(pw, F32), pw.numel(), dptr(pw_out, BF16), dptr(dw, F32), dw.shape[0], dptr(dw_out, F32),
                            cur_stream()), "fd_sep_pack")


def stem_fwd(x, w, bias, y, stride, pad, x_cache=None):
    B, Cin, Hin, Win = x.shape
    C, _, K, _ = w.shape
    is_u8 = 1 if x.dtype == U8 else 0
    check(lib().fd_stem_fwd(dptr(x, U8 if is_u8 else F32), is_u8, dptr(w, F32), dptr(bias, F32), B, Cin, Hin, Win, C, K,
                            stride, pad, dptr(y, BF16), dptr(x_cache, BF16), cur_stream()), "fd_stem_fwd")


def stem_wgrad(x, g, dw, dbias, stride, pad, x_cache=None):
    B, Cin, Hin, Win = x.shape
    C, _, K, _ = dw.shape
    is_u8 = 1 if x.dtype == U8 else 0
    check(lib().fd_stem_wgrad(dptr(x, U8 if is_u8 else F32), is_u8, dptr(g, BF16), B, Cin, Hin, Win, C, K, stride, pad,
                              dptr(dw, F32), dptr(dbias, F32), dptr(x_cache, BF16), cur_stream()), "fd_stem_wgrad")


def stem_fwd_cached(x_cache, in_shape, w, bias, y, stride, pad):
    """Forward of one 64-channel plane from the bf16 image copy a stem_fwd(..., x_cache=) of the same images wrote."""
    B, Cin, Hin, Win = in_shape
    K = w.shape[2]
    assert w.shape[0] == 64
    check(lib().fd_stem_fwd_cached(dptr(x_cache, BF16), dptr(w, F32), dptr(bias, F32), B, Cin, Hin, Win, K, stride, pad,
                                   dptr(y, BF16), cur_stream()), "fd_stem_fwd_cached")


def stem_wgrad_pair(x_cache, in_shape, g0, g1, dw, dbias, stride, pad):
    """dw[128][Cin][K][K], dbias[128] += the stem weight gradients of the two planes g0, g1, one pass over the copy."""
    B, Cin, Hin, Win = in_shape
    K = dw.shape[2]
    assert dw.shape[0] == 128 and dbias.shape[0] == 128
    check(lib().fd_stem_wgrad_pair(dptr(x_cache, BF16), dptr(g0, BF16), dptr(g1, BF16), B, Cin, Hin, Win, K, stride, pad,
                                   dptr(dw, F32), dptr(dbias, F32), cur_stream()), "fd_stem_wgrad_pair")


def stem_cache(B, in_shape3, stride_k_pad, device):
    """Zero-initialised bf16 image copy for the stride-8 tensor-core stem, or None for shapes that do not use one."""
    Cin, Hin, Win = in_shape3
    K, stride, pad = stride_k_pad
    n = stem_cache_elems(B, Cin, Hin, Win, 64, K, stride, pad)
    return torch.zeros(n, dtype=BF16, device=device) if n else None


def stem_planes_fwd(x, w, bias, planes, stride, pad, x_cache=None):
    """Stem of a model with 64 * len(planes) filters: plane 0 from the images (filling x_cache), the others from the copy."""
    for g, y in enumerate(planes):
        wg, bg = w[g * 64:(g + 1) * 64], bias[g * 64:(g + 1) * 64]
        if g > 0 and x_cache is not None:
            stem_fwd_cached(x_cache, x.shape, wg, bg, y, stride, pad)
        else:
            stem_fwd(x, wg, bg, y, stride, pad, x_cache=x_cache if g == 0 else None)


def stem_planes_wgrad(x, g_planes, dw, dbias, stride, pad, x_cache=None):
    """dw[64 G][Cin][K][K], dbias[64 G] += stem weight gradients of the G gradient planes (two planes per pass with a copy)."""
    G = len(g_planes)
    g = 0
    while g < G:
        if g + 1 < G and x_cache is not None:
            stem_wgrad_pair(x_cache, x.shape, g_planes[g], g_planes[g + 1], dw[g * 64:(g + 2) * 64], dbias[g * 64:(g + 2) * 64],
                            stride, pad)
            g += 2
        else:
            stem_wgrad(x, g_planes[g], dw[g * 64:(g + 1) * 64], dbias[g * 64:(g + 1) * 64], stride, pad, x_cache=x_cache)
            g += 1


def head_pack(w, w_t):
    C, K = w.shape[1], w.shape[2]
    check(lib().fd_head_pack(dptr(w, F32), C, K, dptr(w_t, F32), cur_stream()), "fd_head_pack")


def head_fwd(x, chan_scale, w, bias, y, pad, w_t=None):
    B, H, W, C = x.shape
    K = w.shape[2]
    check(lib().fd_head_fwd(dptr(x, BF16), dptr(chan_scale, F32), dptr(w, F32), dptr(w_t, F32), dptr(bias, F32), B, H, W,
                            C, K, pad, dptr(y, F32), cur_stream()), "fd_head_fwd")


def head_bwd(x, chan_scale, w, y, dy, pad, dx, mask_bits, chan_scale2, slope, dx2, dw, dbias, w_t=None):
    B, H, W, C = x.shape
    K = w.shape[2]
    check(lib().fd_head_bwd(dptr(x, BF16), dptr(chan_scale, F32), dptr(w, F32), dptr(w_t, F32), dptr(y, F32),
                            dptr(dy, F32), B, H, W, C,
                            K, pad, dptr(dx, BF16), dptr(mask_bits, I32), dptr(chan_scale2, F32), slope,
                            dptr(dx2, BF16), dptr(dw, F32), dptr(dbias, F32), cur_stream()), "fd_head_bwd")


def maxpool2x2_fwd(x, y, argmax=None):
    """argmax: int16 [B,H/2,W/2,C/8] (2 bits per channel) filled for the backward, or None."""
    B, H, W, C = x.shape
    check(lib().fd_maxpool2x2_fwd(dptr(x, BF16), B, H, W, C, dptr(y, BF16), dptr(argmax, torch.int16), cur_stream()),
          "fd_maxpool2x2_fwd")


def maxpool2x2_bwd(x, gy, gs, mask_bits, chan_scale, slope, gs2, argmax=None):
    """With ``argmax`` (written by maxpool2x2_fwd) the pre-pool tensor ``x`` is not read (only its shape is used)."""
    B, H, W, C = x.shape
    check(lib().fd_maxpool2x2_bwd(None if argmax is not None else dptr(x, BF16), dptr(gy, BF16), B, H, W, C,
                                  dptr(gs, BF16), dptr(mask_bits, I32), dptr(chan_scale, F32), slope, dptr(gs2, BF16),
                                  dptr(argmax, torch.int16), cur_stream()), "fd_maxpool2x2_bwd")


def yolo_loss(pred, gt, loss, dloss_scale=None, dpred=None):
    B, _, S1, S2 = pred.shape
    check(lib().fd_yolo_loss(dptr(pred, F32), dptr(gt, F32), B, S1, S2, dptr(loss, F32), dptr(dloss_scale, F32),
                             dptr(dpred, F32), cur_stream()), "fd_yolo_loss")


def decode_nms(pred, p_thr, iou_thr, width, height, num_of_patches, out_boxes, out_cell, out_count):
    B, _, S1, S2 = pred.shape
    check(lib().fd_decode_nms(dptr(pred, F32), B, S1, S2, float(p_thr), float(iou_thr), int(width), int(height),
                              int(num_of_patches), dptr(out_boxes, F32), dptr(out_cell, I32), dptr(out_count, I32),
                              cur_stream()), "fd_decode_nms")


def box_metrics(gt_boxes, gt_count, pred_boxes, pred_count, iou_thr, out):
    B, cap, _ = gt_boxes.shape
    check(lib().fd_box_metrics(dptr(gt_boxes, F32), dptr(gt_count, I32), dptr(pred_boxes, F32), dptr(pred_count, I32),
                               B, cap, float(iou_thr), dptr(out, F32), cur_stream()), "fd_box_metrics")


def grid_encode(boxes, offsets, S, width, height, out):
    B = out.shape[0]
    check(lib().fd_grid_encode(dptr(boxes, F32), dptr(offsets, I32), B, S, int(width), int(height), dptr(out, F32),
                               cur_stream()), "fd_grid_encode")


def _int_array(vals):
    return (ctypes.c_int * len(vals))(*[int(v) for v in vals])


def ssd_grid_encode(boxes, offsets, patch_sizes, width, height, out):
    B = out.shape[0]
    check(lib().fd_ssd_grid_encode(dptr(boxes, F32), dptr(offsets, I32), B, _int_array(patch_sizes), len(patch_sizes),
                                   int(width), int(height), dptr(out, F32), cur_stream()), "fd_ssd_grid_encode")


def ssd_decode_nms(x, patch_sizes, p_thr, iou_thr, width, height, with_priors, out_boxes, out_count):
    B = x.shape[0]
    check(lib().fd_ssd_decode_nms(dptr(x, F32), B, _int_array(patch_sizes), len(patch_sizes), float(p_thr),
                                  float(iou_thr), int(width), int(height), int(bool(with_priors)),
                                  dptr(out_boxes, F32), dptr(out_count, I32), cur_stream()), "fd_ssd_decode_nms")


def ssd_loss(conf, loc, labels, gt_loc, neg_pos_ratio, row_sums, num_pos, mask=None, dconf=None, dloc=None):
    B, P = conf.shape
    check(lib().fd_ssd_loss(dptr(conf, F32), dptr(loc, F32), dptr(labels, F32), dptr(gt_loc, F32), B, P,
                            int(neg_pos_ratio), dptr(row_sums, F32), dptr(num_pos, I32), dptr(mask, U8),
                            dptr(dconf, F32), dptr(dloc, F32), cur_stream()), "fd_ssd_loss")


# ---------------------------------------------------------------------------------------------- MobilenetV3 path
ACT_NONE, ACT_RELU, ACT_HSWISH = 0, 1, 2


def pw_packed_elems(N, K):
    return int(lib().fd_pw_packed_elems(int(N), int(K)))


def pw_padded_n(N):
    return int(lib().fd_pw_padded_n(int(N)))


def pw_pack(w, scale, out):
    """w [N,K] fp32 (* scale [N], folded BatchNorm) -> packed bf16 operand tiles of fd_pw_conv."""
    N, K = w.shape
    check(lib().fd_pw_pack(dptr(w, F32), dptr(scale, F32), N, K, dptr(out, BF16), cur_stream()), "fd_pw_pack")


def pw_conv(x, w_packed, bias_padded, N, act, out, residual=None):
    """x [..., K] bf16 NHWC -> out [..., N] bf16: 1x1 convolution + bias + act (+ residual) as a tcgen05 GEMM."""
    K = x.shape[-1]
    M = x.numel() // K
    assert out.numel() == M * N and (residual is None or residual.numel() == M * N)
    check(lib().fd_pw_conv(dptr(x, BF16), dptr(w_packed, BF16), dptr(bias_padded, F32), M, K, int(N), int(act),
                           dptr(residual, BF16), dptr(out, BF16), cur_stream()), "fd_pw_conv")


def mbv3_stem(x, w, bias, pad_t, pad_l, out):
    B, _, H, W = x.shape
    _, Ho, Wo, _ = out.shape
    is_u8 = 1 if x.dtype == U8 else 0
    check(lib().fd_mbv3_stem(dptr(x, U8 if is_u8 else F32), is_u8, dptr(w, F32), dptr(bias, F32), B, H, W, int(pad_t),
                             int(pad_l), Ho, Wo, dptr(out, BF16), cur_stream()), "fd_mbv3_stem")


def dw_pack(w, scale, out):
    C, _, K, _ = w.shape
    check(lib().fd_dw_pack(dptr(w, F32), dptr(scale, F32), C, K, dptr(out, F32), cur_stream()), "fd_dw_pack")


def dwconv_se_blocks(Ho, Wo, C):
    return int(lib().fd_dwconv_se_blocks(int(Ho), int(Wo), int(C)))


def dwconv(x, w_packed, bias, K, stride, pad_t, pad_l, act, out, se_partial=None):
    """se_partial: fp32 [B, dwconv_se_blocks(Ho, Wo, C), C] -- per-block channel sums of the output (SqueezeExcite)."""
    B, H, W, C = x.shape
    _, Ho, Wo, _ = out.shape
    assert se_partial is None or tuple(se_partial.shape) == (B, dwconv_se_blocks(Ho, Wo, C), C)
    check(lib().fd_dwconv(dptr(x, BF16), dptr(w_packed, F32), dptr(bias, F32), B, H, W, C, int(K), int(stride), int(pad_t),
                          int(pad_l), Ho, Wo, int(act), dptr(out, BF16), dptr(se_partial, F32), cur_stream()), "fd_dwconv")


def se_gate(se_partial, HW, w1, b1, w2, b2, gate):
    B, nblk, C = se_partial.shape
    R = w1.shape[0]
    check(lib().fd_se_gate(dptr(se_partial, F32), nblk, B, int(HW), dptr(w1, F32), dptr(b1, F32), dptr(w2, F32),
                           dptr(b2, F32), C, R, dptr(gate, F32), cur_stream()), "fd_se_gate")


def scale_channels(x, gate):
    B, H, W, C = x.shape
    check(lib().fd_scale_channels(dptr(x, BF16), dptr(gate, F32), B, H * W, C, cur_stream()), "fd_scale_channels")


def head3x3_fwd(x, w, bias, y):
    B, H, W, C = x.shape
    check(lib().fd_head3x3_fwd(dptr(x, BF16), dptr(w, F32), dptr(bias, F32), B, H, W, C, dptr(y, F32), cur_stream()),
          "fd_head3x3_fwd")


def resize_bilinear(x, out):
    """x [B,C,h,w] -> out [B,C,H,W], both uint8 or both fp32 (transforms.Resize, bilinear, no antialias)."""
    assert x.dtype == out.dtype and x.dtype in (U8, F32)
    B, C, h, w = x.shape
    H, W = out.shape[-2:]
    is_u8 = 1 if x.dtype == U8 else 0
    check(lib().fd_resize_bilinear(dptr(x, x.dtype), is_u8, B * C, h, w, H, W, dptr(out, out.dtype), cur_stream()),
          "fd_resize_bilinear")


def index_copy(dst, src, idx, scatter):
    """scatter: dst[idx[i]] = src[i]; gather: dst[i] = src[idx[i]]  (flat fp32 buffers, int32 index)."""
    check(lib().fd_index_copy_f32(dptr(dst, F32), dptr(src, F32), dptr(idx, I32), idx.numel(), int(bool(scatter)),
                                  cur_stream()), "fd_index_copy_f32")


def _ptr_array(tensors, dtype):
    return (ctypes.c_void_p * len(tensors))(*[dptr(t, dtype) for t in tensors])


def ssd_head_fwd(x_planes, w, bias, mult, priors, prior_off, out):
    """x_planes: list of [B,H,W,64] bf16 planes of one scale; out [B,P,5] fp32 (rows prior_off .. prior_off + H*W)."""
    B, H, W, _ = x_planes[0].shape
    C = w.shape[1]
    check(lib().fd_ssd_head_fwd(_ptr_array(x_planes, BF16), len(x_planes), dptr(w, F32), dptr(bias, F32), B, H * W, C,
                                dptr(mult, F32), dptr(priors, F32), int(prior_off), out.shape[1], dptr(out, F32),
                                cur_stream()), "fd_ssd_head_fwd")


def ssd_head_bwd(x_planes, dx_planes, w, mult, prior_off, out, dout, dw, db):
    B, H, W, _ = x_planes[0].shape
    C = w.shape[1]
    check(lib().fd_ssd_head_bwd(_ptr_array(x_planes, BF16), _ptr_array(dx_planes, BF16), len(x_planes), dptr(w, F32), B,
                                H * W, C, dptr(mult, F32), int(prior_off), out.shape[1], dptr(out, F32), dptr(dout, F32),
                                dptr(dw, F32), dptr(db, F32), cur_stream()), "fd_ssd_head_bwd")


def lrelu_bwd(g, ref, slope, out):
    """out = g * (ref >= +0 ? 1 : slope)  (bf16, same shape)."""
    check(lib().fd_lrelu_bwd(dptr(g, BF16), dptr(ref, BF16), g.numel(), float(slope), dptr(out, BF16), cur_stream()),
          "fd_lrelu_bwd")


def dwconv3x3_wgrad(x, g, dw):
    """dw [64,1,3,3] fp32 += depthwise 3x3 weight gradient of one 64-channel plane (x, g: [B,H,W,64] bf16)."""
    B, H, W, C = x.shape
    check(lib().fd_dwconv3x3_wgrad(dptr(x, BF16), dptr(g, BF16), B, H, W, C, dptr(dw, F32), cur_stream()),
          "fd_dwconv3x3_wgrad")

"""Torch-tensor wrappers over the stage-level C-ABI operators (unit parity tests, ncu captures, bench).

Activations are NHWC matrices `[rows, C]` in torch.float32 or torch.bfloat16, in the FLAT (images*H*W rows) or PADDED
(images*(H+1)*(W+1) rows) row order described in include/viewfusion_b200.h.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return _lib.VF_BF16
    if t.dtype == torch.float32:
        return _lib.VF_F32
    raise TypeError(t.dtype)


def to_nhwc(x: torch.Tensor, dtype=None) -> torch.Tensor:
    """(R,C,H,W) -> FLAT [R*H*W, C] contiguous."""
    y = x.permute(0, 2, 3, 1).contiguous().view(-1, x.shape[1])
    return y if dtype is None else y.to(dtype)


def from_nhwc(y: torch.Tensor, R: int, H: int, W: int) -> torch.Tensor:
    return y.float().view(R, H, W, -1).permute(0, 3, 1, 2).contiguous()


def to_padded(x: torch.Tensor, dtype=None, fill: float = 0.0) -> torch.Tensor:
    """(R,C,H,W) -> PADDED [R*(H+1)*(W+1), C] (host-side reference construction; `fill` goes into the padding rows)."""
    R, Cc, H, W = x.shape
    y = torch.full((R, H + 1, W + 1, Cc), fill, dtype=x.dtype, device=x.device)
    y[:, 1:, 1:, :] = x.permute(0, 2, 3, 1)
    y = y.view(-1, Cc).contiguous()
    return y if dtype is None else y.to(dtype)


def from_padded(y: torch.Tensor, R: int, H: int, W: int) -> torch.Tensor:
    return y.float().view(R, H + 1, W + 1, -1)[:, 1:, 1:, :].permute(0, 3, 1, 2).contiguous()


def flat_to_padded(src: torch.Tensor, images: int, H: int, W: int) -> torch.Tensor:
    lib = _lib.require_device()
    dst = torch.empty(images * (H + 1) * (W + 1), src.shape[1], dtype=src.dtype, device=src.device)
    _lib.check(lib.vf_flat_to_padded(src.data_ptr(), _dt(src), images, H, W, src.shape[1], dst.data_ptr(), _lib.stream_handle()), "vf_flat_to_padded")
    return dst


def padded_to_flat(src: torch.Tensor, images: int, H: int, W: int) -> torch.Tensor:
    lib = _lib.require_device()
    dst = torch.empty(images * H * W, src.shape[1], dtype=src.dtype, device=src.device)
    _lib.check(lib.vf_padded_to_flat(src.data_ptr(), _dt(src), images, H, W, src.shape[1], dst.data_ptr(), _lib.stream_handle()), "vf_padded_to_flat")
    return dst


def pack_conv_weight(w: torch.Tensor, dtype: torch.dtype, cout_pad=None, k_total=None, k_off=0, dst=None) -> torch.Tensor:
    lib = _lib.require_device()
    cout, cin, k, _ = w.shape
    cout_pad = cout if cout_pad is None else cout_pad
    k_total = k * k * cin if k_total is None else k_total
    if dst is None:
        dst = torch.zeros(cout_pad, k_total, dtype=dtype, device=w.device)
    _lib.check(lib.vf_pack_conv_weight(w.contiguous().data_ptr(), cout, cin, k, _dt(dst), dst.data_ptr(), cout_pad, k_total, k_off,
                                       _lib.stream_handle()), "vf_pack_conv_weight")
    return dst


def conv2d(srcs, ksizes, weight, images, H, W, cout, *, stride=1, bias=None, emb=None, img_row=None, residual=None,
           out_dtype=None, out_ld=None, qkv_split=0, cout_pad=None, want_stats=False, in_padded=True, out_padded=True,
           out=None, stats=None):
    """srcs: list of [rows, C] matrices at resolution (H, W); weight: packed [cout_pad, K_total].
    Returns out (+ V^T when qkv_split, + stats when want_stats).  Padding rows of a PADDED output are left as
    allocated (NaN-filled here so that tests catch any read of them)."""
    lib = _lib.require_device()
    a = _lib.ConvArgs()
    dt = srcs[0].dtype
    a.dtype, a.images, a.H, a.W, a.n_seg = _dt(srcs[0]), images, H, W, len(srcs)
    a.in_padded, a.out_padded = int(in_padded), int(out_padded)
    for i, (s, k) in enumerate(zip(srcs, ksizes)):
        a.src[i], a.src_c[i], a.ksize[i] = s.data_ptr(), s.shape[1], k
    a.stride = stride
    a.weight, a.cout, a.cout_pad = weight.data_ptr(), cout, weight.shape[0] if cout_pad is None else cout_pad
    a.bias = _lib.ptr(bias)
    if emb is not None:
        a.emb, a.img_row, a.emb_ld = emb.data_ptr(), img_row.data_ptr(), emb.shape[1]
    a.residual = _lib.ptr(residual)
    out_dtype = dt if out_dtype is None else out_dtype
    out_ld = cout if out_ld is None else out_ld
    Ho, Wo = H // stride, W // stride
    rows = images * ((Ho + 1) * (Wo + 1) if out_padded else Ho * Wo)
    if out is None:
        out = torch.full((rows, out_ld), float("nan"), dtype=out_dtype, device=srcs[0].device)
    a.out, a.out_dtype, a.out_ld = out.data_ptr(), (_lib.VF_BF16 if out_dtype == torch.bfloat16 else _lib.VF_F32), out_ld
    vt = None
    if qkv_split and dt == torch.bfloat16:
        vt = torch.zeros(images, qkv_split, H * W, dtype=dt, device=out.device)
        a.qkv_split, a.out_vt = qkv_split, vt.data_ptr()
    if want_stats:
        if stats is None:
            stats = torch.zeros(images, cout, 2, dtype=torch.float32, device=out.device)
        a.stats = stats.data_ptr()
    _lib.check(lib.vf_conv2d(C.byref(a), _lib.stream_handle()), "vf_conv2d")
    if want_stats:
        return out, stats
    return (out, vt) if qkv_split else out


def gn_stats(src0, src1, images, H, W):
    lib = _lib.require_device()
    C0, C1 = src0.shape[1], (0 if src1 is None else src1.shape[1])
    stats = torch.zeros(images, C0 + C1, 2, dtype=torch.float32, device=src0.device)
    _lib.check(lib.vf_gn_stats(src0.data_ptr(), C0, _lib.ptr(src1), C1, _dt(src0), images, H, W, stats.data_ptr(), _lib.stream_handle()),
               "vf_gn_stats")
    return stats


def gn_shift(bias=None, emb=None, img_row=None):
    """vf_gn_shift for source 0: stored without bias[c] + emb[img_row[img]][c] (keep the tensors alive during the call)."""
    s = _lib.GnShift()
    s.bias, s.emb, s.img_row = _lib.ptr(bias), _lib.ptr(emb), _lib.ptr(img_row)
    s.emb_ld = 0 if emb is None else emb.shape[1]
    return s


def gn_apply(src0, src1, images, H, W, groups, stats, gamma, beta, swish=True, stats1=None, shift=None):
    """PADDED in, PADDED out.  stats: [images, C0+C1, 2] (from gn_stats), or per-source [images, C0, 2] + stats1."""
    lib = _lib.require_device()
    C0, C1 = src0.shape[1], (0 if src1 is None else src1.shape[1])
    dst = torch.full((images * (H + 1) * (W + 1), C0 + C1), float("nan"), dtype=src0.dtype, device=src0.device)
    if stats1 is None:
        s0, ld0, s1, ld1 = stats.data_ptr(), stats.shape[1], stats.data_ptr() + 8 * C0, stats.shape[1]
    else:
        s0, ld0, s1, ld1 = stats.data_ptr(), stats.shape[1], stats1.data_ptr(), stats1.shape[1]
    _lib.check(lib.vf_gn_apply(src0.data_ptr(), C0, s0, ld0, _lib.ptr(src1), C1, s1 if C1 else 0, ld1, _dt(src0), images, H, W, groups,
                               gamma.data_ptr(), beta.data_ptr(), int(swish), dst.data_ptr(), None if shift is None else C.byref(shift),
                               _lib.stream_handle()), "vf_gn_apply")
    return dst


def upsample2x(src, images, H, W):
    lib = _lib.require_device()
    Cc = src.shape[1]
    dst = torch.full((images * (2 * H + 1) * (2 * W + 1), Cc), float("nan"), dtype=src.dtype, device=src.device)
    _lib.check(lib.vf_upsample2x(src.data_ptr(), _dt(src), images, H, W, Cc, dst.data_ptr(), _lib.stream_handle()), "vf_upsample2x")
    return dst


def attention(qkv, vt, images, L, Cc, lse=None):
    lib = _lib.require_device()
    out = torch.empty(images * L, Cc, dtype=qkv.dtype, device=qkv.device)
    _lib.check(lib.vf_attention(qkv.data_ptr(), _lib.ptr(vt), _dt(qkv), images, L, Cc, out.data_ptr(), _lib.ptr(lse), _lib.stream_handle()),
               "vf_attention")
    return out


def attention_backward(qkv, vt, out, lse, d_out, images, L, Cc):
    """dqkv [images*L, 3C] of softmax(QK^T/sqrt(C))V given d_out (tensor cores when out/lse are given and the shape fits)."""
    lib = _lib.require_device()
    scratch = torch.empty(images * L * 2 * Cc * max(1, L // 128), dtype=torch.float32, device=qkv.device)
    dqkv = torch.empty(images * L, 3 * Cc, dtype=qkv.dtype, device=qkv.device)
    _lib.check(lib.vf_attention_backward(qkv.data_ptr(), _lib.ptr(vt), _lib.ptr(out), _lib.ptr(lse), d_out.data_ptr(), _dt(qkv), images, L, Cc,
                                         scratch.data_ptr(), dqkv.data_ptr(), _lib.stream_handle()), "vf_attention_backward")
    return dqkv


def embed(level, angle, ic, w0, b0, w2, b2, emb_w, emb_b):
    lib = _lib.require_device()
    rows, E = level.numel(), emb_w.shape[0]
    out = torch.empty(rows, E, dtype=torch.float32, device=level.device)
    _lib.check(lib.vf_embed(level.data_ptr(), angle.data_ptr(), rows, ic, w0.data_ptr(), b0.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                            emb_w.data_ptr(), emb_b.data_ptr(), E, out.data_ptr(), _lib.stream_handle()), "vf_embed")
    return out


def force_simt(on: bool) -> None:
    _lib.load().vf_debug_force_simt(int(on))

"""Input path on the device: `prepare_batch` turns the decoded uint8 views of a batch of objects into
the (target, cond, angle) tensors the reference's loader yields per sample (data/nmr_dataset.py:10-52), in one launch.
Only the uint8 bytes cross PCIe (4x less than the float images the reference collates on the host)."""
from __future__ import annotations

import torch

from . import _lib


def prepare_batch(views_u8: torch.Tensor, perm: torch.Tensor):
    """views_u8 (B, V, H, W, C) uint8 on the GPU, perm (B, V) integer view order per object (perm[b, 0] = target view).
    Returns target (B, C, H, W), cond (B, V-1, C, H, W), angle (B, 1), fp32 on the same device."""
    lib = _lib.require_device()
    if views_u8.dtype != torch.uint8 or views_u8.dim() != 5:
        raise ValueError(f"views_u8 must be a (B, V, H, W, C) uint8 tensor, got {views_u8.dtype} {tuple(views_u8.shape)}")
    if not views_u8.is_cuda:
        raise RuntimeError("view_fusion_b200.inputs needs CUDA tensors; there is no CPU fallback")
    B, V, H, W, C = views_u8.shape
    if tuple(perm.shape) != (B, V):
        raise ValueError(f"perm must be ({B}, {V}), got {tuple(perm.shape)}")
    p32 = perm.to(device=views_u8.device, dtype=torch.int32).contiguous()
    if int(p32.min()) < 0 or int(p32.max()) >= V:
        raise ValueError("perm entries outside [0, V)")
    v = views_u8.contiguous()
    target = torch.empty(B, C, H, W, dtype=torch.float32, device=v.device)
    cond = torch.empty(B, V - 1, C, H, W, dtype=torch.float32, device=v.device)
    angle = torch.empty(B, 1, dtype=torch.float32, device=v.device)
    _lib.check(lib.vf_prepare_batch_u8(v.data_ptr(), p32.data_ptr(), B, V, C, H, W, target.data_ptr(), cond.data_ptr(), angle.data_ptr(),
                                       _lib.stream_handle()), "vf_prepare_batch_u8")
    return target, cond, angle

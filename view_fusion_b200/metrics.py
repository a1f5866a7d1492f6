"""Evaluation metrics on the device: drop-ins for `utils/metrics.py:6-12`.

`compute_psnr(generated, target)` and `compute_ssim(generated, target)` take the reference's (B, C, H, W) fp32 tensors in
[0, 1] and return (B,) tensors, like the reference; both come out of ONE kernel launch (`vf_eval_metrics`), so asking for
both through `compute_psnr_ssim` costs one pass.  There is no CPU fallback.
"""
from __future__ import annotations

import torch

from . import _lib


def compute_psnr_ssim(generated: torch.Tensor, target: torch.Tensor):
    lib = _lib.require_device()
    if generated.shape != target.shape or generated.dim() != 4:
        raise ValueError(f"expected two (B, C, H, W) tensors, got {tuple(generated.shape)} and {tuple(target.shape)}")
    if not (generated.is_cuda and target.is_cuda):
        raise RuntimeError("view_fusion_b200.metrics needs CUDA tensors; there is no CPU fallback")
    g, t = generated.contiguous().float(), target.contiguous().float()
    B, C, H, W = g.shape
    psnr = torch.empty(B, dtype=torch.float32, device=g.device)
    ssim = torch.empty(B, dtype=torch.float32, device=g.device)
    _lib.check(lib.vf_eval_metrics(g.data_ptr(), t.data_ptr(), B, C, H, W, psnr.data_ptr(), ssim.data_ptr(), _lib.stream_handle()),
               "vf_eval_metrics")
    return psnr, ssim


def compute_psnr(generated: torch.Tensor, target: torch.Tensor) -> torch.Tensor:      # utils/metrics.py:6-8
    return compute_psnr_ssim(generated, target)[0]


def compute_ssim(generated: torch.Tensor, target: torch.Tensor) -> torch.Tensor:      # utils/metrics.py:11-12
    return compute_psnr_ssim(generated, target)[1]

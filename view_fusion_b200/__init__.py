"""view_fusion_b200 — B200-native (sm_100a) implementation of ViewFusion's data-parallel hot path.

Public surface (mirrors the reference's `model/` package):
    UNet        — drop-in for model/unet.py:UNet
    ViewFusion  — drop-in for model/view_fusion.py:ViewFusion
The arithmetic lives in `libviewfusion_b200.so` (hand-written CUDA behind the C ABI of include/viewfusion_b200.h).
"""
from .unet import UNet
from .view_fusion import ViewFusion, make_beta_schedule

__all__ = ["UNet", "ViewFusion", "make_beta_schedule"]

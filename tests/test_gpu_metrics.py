"""vf_eval_metrics (PSNR + SSIM on the device) against the oracle restatement of utils/metrics.py:6-12."""
import pytest
import torch

import vf_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,C,H,W", [(28, 3, 64, 64), (5, 3, 32, 48), (2, 1, 11, 11), (3, 3, 80, 88)])
def test_psnr_ssim_match_oracle(B, C, H, W):
    from view_fusion_b200 import metrics
    g = torch.Generator().manual_seed(B * 100 + H)
    target = torch.rand(B, C, H, W, generator=g)
    noise = torch.randn(B, C, H, W, generator=g) * torch.linspace(0.01, 0.3, B).view(B, 1, 1, 1)
    generated = (target + noise).clamp(0, 1)
    psnr, ssim = metrics.compute_psnr_ssim(generated.cuda(), target.cuda())
    torch.cuda.synchronize()
    assert psnr.shape == ssim.shape == (B,)
    assert torch.allclose(psnr.cpu(), O.psnr(generated, target), rtol=1e-5, atol=1e-4)
    assert torch.allclose(ssim.cpu(), O.ssim(generated, target), rtol=1e-4, atol=2e-5)
    # identical images: SSIM exactly 1 within rounding, PSNR infinite like the reference's formula
    p1, s1 = metrics.compute_psnr_ssim(target.cuda(), target.cuda())
    assert torch.allclose(s1.cpu(), torch.ones(B), atol=1e-5) and bool(torch.isinf(p1).all())


def test_metrics_reject_bad_shapes():
    from view_fusion_b200 import metrics
    with pytest.raises(RuntimeError):
        metrics.compute_psnr_ssim(torch.rand(1, 3, 8, 8).cuda(), torch.rand(1, 3, 8, 8).cuda())       # smaller than the window
    with pytest.raises(ValueError):
        metrics.compute_psnr_ssim(torch.rand(1, 3, 16, 16).cuda(), torch.rand(1, 3, 16, 17).cuda())
    with pytest.raises(RuntimeError):
        metrics.compute_psnr_ssim(torch.rand(1, 3, 16, 16), torch.rand(1, 3, 16, 16))

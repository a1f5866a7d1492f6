"""CPU: drop-in contract of the Python modules (constructor, state_dict, error behaviour) — no GPU compute."""
import contextlib
import io

import numpy as np
import pytest
import torch

import vf_oracle as O
from view_fusion_b200 import UNet, ViewFusion, make_beta_schedule

BETA = {"train": dict(O.BETA_TRAIN)}


def _vf(cfg, **kw):
    with contextlib.redirect_stdout(io.StringIO()) as out:
        m = ViewFusion(UNet(**cfg), BETA, **kw)
    return m, out.getvalue()


def test_state_dict_layout_matches_reference():
    m, printed = _vf(O.SMALL_V100)
    assert printed.startswith("Weighting train and inference: True True")          # view_fusion.py:29-33
    m.set_new_noise_schedule(device="cpu", phase="train")
    sd = m.state_dict()
    want = ["gammas", "sqrt_recip_gammas", "sqrt_recipm1_gammas", "posterior_log_variance_clipped",
            "posterior_mean_coef1", "posterior_mean_coef2"] + ["denoise_fn." + n for n, _, _ in O.param_shapes(O.SMALL_V100)]
    assert list(sd) == want and len(sd) == 406
    for n, shp, _ in O.param_shapes(O.SMALL_V100):
        assert tuple(sd["denoise_fn." + n].shape) == tuple(shp)
    assert sum(p.numel() for p in m.denoise_fn.parameters()) == 33_947_206
    assert m.num_timesteps == 2000


def test_schedule_buffers_bit_exact(golden_dir):
    import os
    g = np.load(os.path.join(golden_dir, "schedule_train.npz"))
    m, _ = _vf(O.TINY)
    m.set_new_noise_schedule(device="cpu", phase="train")
    for k in g.files:
        assert np.array_equal(getattr(m, k).numpy(), g[k]), k


def test_state_dict_roundtrip_with_oracle_weights():
    m, _ = _vf(O.TINY)
    m.set_new_noise_schedule(device="cpu", phase="train")
    sd = O.init_state_dict(O.TINY, 5, prefix="denoise_fn.")
    sd.update(O.make_schedule(**O.BETA_TRAIN))
    m.load_state_dict(sd, strict=True)
    back = m.state_dict()
    assert all(torch.equal(back[k], sd[k]) for k in sd)


def test_reference_error_conventions():
    with pytest.raises(NotImplementedError):
        make_beta_schedule("nope", 10)
    m, _ = _vf(O.TINY)
    m.beta_schedule = {"train": dict(schedule="linear", num_timesteps=4, linear_start=1e-4, linear_end=1e-2)}
    m.set_new_noise_schedule(device="cpu")
    if torch.cuda.is_available():
        pytest.skip("assertion order needs CPU tensors")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.generate(torch.zeros(1, 2, 3, 16, 16), torch.tensor([2]), torch.zeros(1, 1))


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    net = UNet(**O.TINY)
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 6, 16, 16), torch.zeros(1, 1), torch.zeros(1, 1))


def test_no_cpu_fallback_in_metrics_and_inputs():
    """The eval metrics and the input path are CUDA-only too: on a machine without a B200 (or with CPU tensors) they raise."""
    from view_fusion_b200 import inputs, metrics
    a = torch.rand(2, 3, 16, 16)
    with pytest.raises(RuntimeError):
        metrics.compute_psnr(a, a)
    with pytest.raises(RuntimeError):
        metrics.compute_ssim(a, a)
    with pytest.raises(RuntimeError):
        inputs.prepare_batch(torch.zeros(1, 2, 4, 4, 3, dtype=torch.uint8), torch.zeros(1, 2, dtype=torch.long))

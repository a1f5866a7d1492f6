"""CPU: the oracle (oracle/vf_oracle.py) against the committed REFERENCE outputs in tests/golden/.

The fixtures were produced by oracle/make_golden.py from the unmodified reference modules; these tests
make sure the restatement still reproduces them wherever the suite runs (the reference itself cannot travel).
"""
import os

import numpy as np
import pytest
import torch

import vf_oracle as O

torch.set_num_threads(max(1, os.cpu_count() or 1))


def _load(golden_dir, name):
    d = np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False)
    return {k: (torch.from_numpy(d[k]) if d[k].dtype.kind in "fi" and d[k].ndim > 0 else d[k]) for k in d.files}


def rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def test_schedule_bit_exact(golden_dir):
    g = _load(golden_dir, "schedule_train")
    s = O.make_schedule(**O.BETA_TRAIN)
    assert set(s) == set(g)
    for k in s:
        assert torch.equal(s[k], g[k]), k
    assert s["gammas"].shape == (2000,)


def test_layout_matches_survey_appendix():
    shapes = O.param_shapes(O.SMALL_V100)
    assert len(shapes) == 400
    assert sum(int(np.prod(s)) for _, s, _ in shapes) == 33_947_206


@pytest.mark.parametrize("tag,cfg", [("tiny", O.TINY), ("small", O.SMALL_V100)])
def test_unet_forward(golden_dir, tag, cfg):
    g = _load(golden_dir, f"unet_{tag}")
    sd = O.init_state_dict(cfg, int(g["seed"]))
    chk = sum(float(v.double().sum()) for v in sd.values())
    assert abs(chk - float(g["w_checksum"])) < 1e-9 * max(1.0, abs(chk)), "weight init is not reproducible on this box"
    with torch.no_grad():
        out = O.unet_forward(sd, cfg, g["x"], g["angle"], g["time"])
    assert rel(out, g["out"]) < 5e-6


@pytest.mark.parametrize("tag,weighting", [("ragged", True), ("full6", True), ("mean", False)])
def test_compose_ddpm(golden_dir, tag, weighting):
    g = _load(golden_dir, f"compose_{tag}")
    sched = O.make_schedule(**O.BETA_TRAIN)
    eps, logits, w = O.compose(g["out"], g["view_count"], weighting)
    assert rel(eps, g["eps"]) < 1e-6
    if weighting:
        assert rel(w, g["weights"]) < 1e-6
        # padded slots carry exactly zero weight, real slots sum to one
        vc = g["view_count"].tolist()
        for b, v in enumerate(vc):
            assert float(w[b, v:].abs().sum()) == 0.0
        assert torch.allclose(w.sum(dim=1), torch.ones_like(w.sum(dim=1)), atol=1e-6)
    for tn in ("hi", "mid", "one", "zero"):
        y = O.ddpm_update(sched, g["y_t"], eps, g[f"t_{tn}"], g["z"])
        assert rel(y, g[f"y_prev_{tn}"]) < 1e-6, tn


def test_psample_tiny_trajectory(golden_dir):
    g = _load(golden_dir, "psample_tiny_ragged")
    cfg = O.TINY
    sd = O.init_state_dict(cfg, int(g["seed"]), prefix="denoise_fn.")
    sched = O.make_schedule(**O.BETA_TRAIN)
    y = g["y_T"]
    with torch.no_grad():
        for j, i in enumerate(g["steps"].tolist()):
            t = torch.full((y.shape[0],), i, dtype=torch.long)
            y, eps, logits, w = O.p_sample(sd, cfg, sched, y, g["y_cond"], g["view_count"], g["angle"], t, g["z"][j])
            assert rel(y, g["y"][j]) < 1e-5, i
    assert rel(w, g["weights_last"]) < 1e-5
    assert rel(logits, g["logits_last"]) < 1e-5


def test_train_tiny_loss_and_grads(golden_dir):
    g = _load(golden_dir, "train_tiny_ragged")
    cfg = O.TINY
    sd = O.init_state_dict(cfg, int(g["seed"]), prefix="denoise_fn.")
    sd = {k: v.requires_grad_(True) for k, v in sd.items()}
    sched = O.make_schedule(**O.BETA_TRAIN)
    loss, eps = O.train_loss(sd, cfg, sched, g["y_0"], g["y_cond"], g["view_count"], g["angle"], g["t"], g["u"], g["noise"])
    loss.backward()
    assert abs(float(loss.detach()) - float(g["loss"])) < 2e-6
    assert rel(eps.detach(), g["eps"]) < 1e-5
    names = [str(n) for n in g["grad_names"]]
    norms = g["grad_norms"]
    big = float(norms.max())
    for n, ref in zip(names, norms.tolist()):
        mine = float(sd["denoise_fn." + n].grad.norm())
        if ref < 1e-6 * big:          # mathematically-zero gradients (bias in front of a 1-channel-per-group GN)
            assert mine < 1e-6 * big, n
        else:
            assert abs(mine - ref) < 1e-4 * ref, n
    for k in g:
        if k.startswith("grad:"):
            full = sd["denoise_fn." + k[5:]].grad.reshape(-1)
            stride = max(1, (full.numel() + 8191) // 8192)
            assert rel(full[::stride], g[k]) < 1e-4, k


def test_generate_contract_shapes():
    """generate() return tuple shapes of view_fusion.py:208-214 on a 16-step toy schedule."""
    cfg = O.TINY
    sd = O.init_state_dict(cfg, 3, prefix="denoise_fn.")
    sched = O.make_schedule(schedule="linear", num_timesteps=16, linear_start=1e-4, linear_end=0.09)
    bt = O.synthetic_batch(2, 3, 16, seed=5, ragged=True)
    zs = O.normal_draws(16, (2, 3, 16, 16), seed=6)
    y_T = O.normal_draws(1, (2, 3, 16, 16), seed=7)[0]
    with torch.no_grad():
        y, ret, la, wa, last = O.generate(sd, cfg, sched, bt["y_cond"], bt["view_count"], bt["angle"], y_T, zs)
    sv, mv = int(bt["view_count"].sum()), int(bt["view_count"].max())
    assert ret.shape == (2, 9, 3, 16, 16) and la.shape == (sv, 8, 3, 16, 16) and wa.shape == (2, 8, mv, 3, 16, 16)
    assert torch.equal(last, y)


def test_ssim_restatement_against_bruteforce_float64():
    """O.ssim (pytorch-msssim 1.0.0's published algorithm; the package is absent here, so parity is unpinned) against a
    direct float64 evaluation of the same definition, plus the properties any SSIM has."""
    import numpy as np
    g = torch.Generator().manual_seed(4)
    a = torch.rand(2, 3, 20, 24, generator=g)
    b = (a + 0.1 * torch.randn(2, 3, 20, 24, generator=g)).clamp(0, 1)
    got = O.ssim(a, b)
    k = np.arange(11) - 5
    w = np.exp(-(k ** 2) / (2 * 1.5 ** 2)); w /= w.sum()
    w2 = np.outer(w, w)
    an, bn = a.double().numpy(), b.double().numpy()
    want = np.zeros(2)
    for n in range(2):
        for c in range(3):
            vals = []
            for i in range(20 - 10):
                for j in range(24 - 10):
                    x, y = an[n, c, i:i + 11, j:j + 11], bn[n, c, i:i + 11, j:j + 11]
                    mx, my = (w2 * x).sum(), (w2 * y).sum()
                    sxx, syy, sxy = (w2 * x * x).sum() - mx * mx, (w2 * y * y).sum() - my * my, (w2 * x * y).sum() - mx * my
                    vals.append((2 * mx * my + 1e-4) / (mx * mx + my * my + 1e-4) * (2 * sxy + 9e-4) / (sxx + syy + 9e-4))
            want[n] += np.mean(vals) / 3
    assert np.allclose(got.numpy(), want, atol=2e-6)
    assert torch.allclose(O.ssim(a, a), torch.ones(2), atol=1e-6)
    assert torch.allclose(O.ssim(a, b), O.ssim(b, a), atol=1e-6)
    assert bool((got < 1).all())


def test_process_batch_restatement_matches_reference_semantics():
    """O.process_batch_u8 against the literal per-sample steps of data/nmr_dataset.py:10-24, 43 with the same permutation."""
    import numpy as np
    rng = np.random.default_rng(0)
    views = rng.integers(0, 256, size=(2, 24, 6, 5, 3), dtype=np.uint8)
    perm = np.stack([rng.permutation(24) for _ in range(2)])
    got = O.process_batch_u8(views, perm)
    for b in range(2):
        images = np.stack([views[b, i].astype("f") / 255.0 for i in range(24)], 0).astype(np.float32)   # decode("rgb") + np.stack
        images = np.transpose(images, (0, 3, 1, 2))
        cond_images = images[perm[b]]
        assert np.array_equal(got["target"][b], cond_images[0])
        assert np.array_equal(got["cond"][b], cond_images[1:])
        assert got["angle"][b, 0] == np.asarray([2 * np.pi / 24 * perm[b, 0]]).astype(np.float32)[0]
    assert got["cond"].shape == (2, 23, 3, 6, 5) and got["target"].dtype == np.float32

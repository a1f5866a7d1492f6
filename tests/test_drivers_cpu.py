"""Host logic of view_fusion_b200.drivers against the reference's calling patterns (experiment.py:472-620), with a stub
in place of the model: the drivers only assemble (y_cond, view_count, angle) and call `model(..., generate=True)`."""
import math

import pytest
import torch

from view_fusion_b200 import drivers


class StubModel:
    """generate=True -> 5-tuple; the 'generated' view is a deterministic function of the live conditioning prefix and the
    angle, so that any mistake in view_count / angle / append order changes the result."""

    def __init__(self):
        self.calls = []

    def __call__(self, y_cond, view_count, angle, generate=False):
        assert generate
        self.calls.append((tuple(y_cond.shape), view_count.clone(), angle.clone()))
        B = y_cond.shape[0]
        out = torch.stack([y_cond[b, : int(view_count[b])].mean(0) * 0.5 + 0.1 * torch.cos(angle[b, 0]) for b in range(B)])
        return out, out[:, None], None, None, out


def _reference_autoregressive(model, cond, n):
    """The loop of experiment.py:516-545 (re-concatenating the conditioning tensor every step)."""
    samples = []
    for count in range(1, n + 1):
        a = 2 * math.pi / n * count
        view_count = torch.full((cond.shape[0],), count)
        angle = torch.full((cond.shape[0], 1), a)
        *_, gen = model(y_cond=cond, view_count=view_count, angle=angle, generate=True)
        cond = torch.cat((cond, gen[:, None, ...]), dim=1)
        samples.append(gen)
    return cond, torch.stack(samples)


def test_autoregressive_orbit_equals_reference_loop():
    torch.manual_seed(0)
    first = torch.rand(2, 1, 3, 8, 8)
    ref_cond, ref_samples = _reference_autoregressive(StubModel(), first.clone(), 6)
    m = StubModel()
    cond, samples = drivers.autoregressive_orbit(m, first, n_targets=6)
    assert torch.equal(cond, ref_cond) and torch.equal(samples, ref_samples)
    assert cond.shape == (2, 7, 3, 8, 8) and samples.shape == (6, 2, 3, 8, 8)
    # one buffer of the final size, the live prefix selected by view_count (no re-allocation while the set grows)
    assert all(shape == (2, 7, 3, 8, 8) for shape, _, _ in m.calls)
    assert [int(vc[0]) for _, vc, _ in m.calls] == [1, 2, 3, 4, 5, 6]
    assert all(vc.device.type == "cpu" and vc.dtype == torch.long for _, vc, _ in m.calls)
    assert torch.allclose(torch.stack([a[0, 0] for _, _, a in m.calls]), torch.tensor([2 * math.pi * k / 6 for k in range(1, 7)]))


def test_autoregressive_orbit_accepts_4d_and_custom_angles_and_clamps():
    first = torch.full((1, 3, 4, 4), 3.0)
    m = StubModel()
    cond, samples = drivers.autoregressive_orbit(m, first, n_targets=2, angles=[0.0, math.pi], clamp=True)
    assert float(cond[:, 1:].max()) <= 1.0 and float(cond[:, 0].max()) == 3.0
    assert [float(a[0, 0]) for _, _, a in m.calls] == [0.0, pytest.approx(math.pi)]
    with pytest.raises(ValueError):
        drivers.autoregressive_orbit(m, first, n_targets=2, angles=[0.0])
    with pytest.raises(ValueError):
        drivers.autoregressive_orbit(m, torch.rand(1, 2, 3, 4, 4), n_targets=2)


def test_extrapolate_draws_view_counts_like_the_reference():
    cond, angle = torch.rand(5, 23, 3, 4, 4), torch.zeros(5, 1)
    g = torch.Generator().manual_seed(7)
    want = torch.randint(7, 24, (5,), generator=torch.Generator().manual_seed(7))      # experiment.py:477: randint(max_views + 1, 24)
    m = StubModel()
    vc, out = drivers.extrapolate(m, cond, angle, min_views=7, max_views=24, generator=g)
    assert torch.equal(vc, want) and len(out) == 5
    assert torch.equal(m.calls[0][1], want)
    vc2, _ = drivers.extrapolate(m, cond, angle, min_views=0, view_count=torch.tensor([1, 23, 5, 5, 9]))
    assert vc2.tolist() == [1, 23, 5, 5, 9]
    with pytest.raises(ValueError):
        drivers.extrapolate(m, cond, angle, min_views=0, view_count=torch.tensor([1, 24, 5, 5, 9]))
    with pytest.raises(ValueError):
        drivers.extrapolate(m, cond, angle, min_views=30, max_views=40)


def test_orbit_from_views_batches_all_target_angles():
    views = torch.rand(24, 3, 4, 4)
    m = StubModel()
    out = drivers.orbit_from_views(m, views, cond_stride=4, n_targets=24)
    shape, vc, angle = m.calls[0]
    assert shape == (24, 6, 3, 4, 4) and vc.tolist() == [6] * 24 and len(out) == 5
    assert torch.allclose(angle[:, 0], torch.tensor([2 * math.pi * k / 24 for k in range(24)]))
    with pytest.raises(TypeError):
        drivers.orbit_from_views(lambda **kw: (1, 2), views)

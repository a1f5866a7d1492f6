"""Sampling drivers around `ViewFusion.generate`: the calling patterns of the reference's visualisation
modes, `experiment.py:472-650`, without its wandb / torchvision glue.

They only build the (y_cond, view_count, angle) arguments the way the reference does and call the model's public API, so
they work with any object that has the reference's `generate=True` call signature; the CUDA work is all behind
`ViewFusion.generate`.

  autoregressive_orbit   experiment.py:516-560 (`-ar`): one primed view per object; target k (angle 2*pi*k/n) is generated
                         from the k views available so far and appended to the conditioning set.  The conditioning buffer
                         is allocated once at its final size (B, n+1, C, H, W) and the generated views are written into
                         it in place — `view_count = k` selects the live prefix, so nothing is re-allocated or copied
                         while the set grows (the reference re-concatenates the whole tensor every step).
  extrapolate            experiment.py:472-514 (`-ex`): more conditioning views than the training maximum, drawn per sample.
  orbit_from_views       experiment.py:582-620 (`-gif`): all n target angles of one object from a fixed subset of its views,
                         as ONE batch (the n targets are independent samples).
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import torch


def _call_generate(model, y_cond, view_count, angle):
    out = model(y_cond=y_cond, view_count=view_count, angle=angle, generate=True)
    if not isinstance(out, (tuple, list)) or len(out) != 5:
        raise TypeError("model(..., generate=True) must return the reference's 5-tuple (view_fusion.py:214)")
    return out


@torch.no_grad()
def autoregressive_orbit(model, first_view: torch.Tensor, n_targets: int = 24, angles: Optional[Sequence[float]] = None,
                         clamp: bool = False):
    """first_view (B, C, H, W) or (B, 1, C, H, W).  Returns (cond, samples):
    cond (B, n_targets + 1, C, H, W) = the primed view followed by every generated view, samples (n_targets, B, C, H, W).
    `angles[k-1]` is the target angle of step k (default 2*pi*k/n_targets, experiment.py:520-522).  `clamp` stores the
    generated views clamped to [0, 1] (the reference feeds them back unclamped)."""
    if first_view.dim() == 4:
        first_view = first_view[:, None]
    if first_view.dim() != 5 or first_view.shape[1] != 1:
        raise ValueError(f"first_view must be (B, C, H, W) or (B, 1, C, H, W), got {tuple(first_view.shape)}")
    if n_targets < 1:
        raise ValueError("n_targets must be positive")
    if angles is None:
        angles = [2.0 * math.pi * k / n_targets for k in range(1, n_targets + 1)]
    if len(angles) != n_targets:
        raise ValueError(f"{len(angles)} angles for {n_targets} targets")
    B, _, C, H, W = first_view.shape
    dev = first_view.device
    cond = torch.zeros(B, n_targets + 1, C, H, W, dtype=first_view.dtype, device=dev)
    cond[:, 0] = first_view[:, 0]
    samples = []
    for count, a in enumerate(angles, start=1):
        view_count = torch.full((B,), count, dtype=torch.long)          # host-resident: no device sync to read it
        angle = torch.full((B, 1), float(a), dtype=torch.float32, device=dev)
        *_, generated = _call_generate(model, cond, view_count, angle)
        generated = generated[:, :C]
        cond[:, count] = generated.clamp(0, 1) if clamp else generated   # in place: the live prefix grows by one view
        samples.append(cond[:, count].clone())
    return cond, torch.stack(samples)


@torch.no_grad()
def extrapolate(model, cond: torch.Tensor, angle: torch.Tensor, min_views: int, max_views: Optional[int] = None,
                view_count: Optional[torch.Tensor] = None, generator: Optional[torch.Generator] = None):
    """cond (B, Nmax, C, H, W), angle (B, 1).  view_count ~ randint(min_views, max_views) per sample unless given
    (experiment.py:477-479 draws it in [max_views_train + 1, 24)).  Returns (view_count, generate()'s 5-tuple)."""
    B, n_max = cond.shape[:2]
    max_views = n_max + 1 if max_views is None else max_views
    if view_count is None:
        if not (1 <= min_views < max_views <= n_max + 1):
            raise ValueError(f"need 1 <= min_views < max_views <= Nmax + 1, got {min_views}, {max_views}, Nmax = {n_max}")
        view_count = torch.randint(min_views, max_views, (B,), generator=generator)
    view_count = view_count.to(torch.long).cpu()
    if int(view_count.max()) > n_max or int(view_count.min()) < 1:
        raise ValueError("view_count outside [1, Nmax]")
    return view_count, _call_generate(model, cond, view_count, angle)


@torch.no_grad()
def orbit_from_views(model, views: torch.Tensor, cond_stride: int = 4, n_targets: int = 24):
    """views (V, C, H, W): all views of ONE object.  Conditions on views[::cond_stride] and generates the n_targets angles
    2*pi*k/n_targets, k = 0..n_targets-1, as one batch (experiment.py:582-600).  Returns generate()'s 5-tuple."""
    if views.dim() != 4:
        raise ValueError(f"views must be (V, C, H, W), got {tuple(views.shape)}")
    sel = views[::cond_stride]
    y_cond = sel[None].expand(n_targets, *sel.shape).contiguous()
    angle = torch.tensor([2.0 * math.pi * k / n_targets for k in range(n_targets)], dtype=torch.float32, device=views.device)[:, None]
    view_count = torch.full((n_targets,), sel.shape[0], dtype=torch.long)
    return _call_generate(model, y_cond, view_count, angle)

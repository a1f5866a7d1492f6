"""GPU: the device-resident reverse loop (vf_p_sample_step + CUDA-graph replay) and the host-side state rules.

  * generate() replays ONE captured CUDA graph per reverse step (time-step, noise level, `any(t > 0)` and the Philox
    offset / seed live on the device): the result must equal the eager loop, and a capture must really have happened;
  * the in-kernel noise follows torch.manual_seed like the reference's randn_like (view_fusion.py:176);
  * workspaces are laid out for a capacity, so a different batch size / view counts at the same (or a smaller) image
    count reuse the zero-filled arena without corrupting the padding-row invariant (fp32 mode: the CUDA-core path);
  * a backward after ANOTHER forward on the same UNet is refused instead of silently using the wrong activation stash.
Reference semantics: model/view_fusion.py:166-214, experiment.py:277, :288-293."""
import math

import pytest
import torch

import vf_oracle as O
from gpu_util import TOY64, build_model, margin, rel

pytestmark = pytest.mark.gpu


def _inputs(B, N, S, seed):
    g = torch.Generator().manual_seed(seed)
    y_cond = torch.rand(B, N, 3, S, S, generator=g)
    y_T = torch.randn(B, 3, S, S, generator=g)
    angle = (2 * math.pi / 24) * torch.randint(0, 24, (B, 1), generator=g).float()
    return y_cond.cuda(), y_T.cuda(), angle.cuda()


@pytest.mark.parametrize("prec,cfg,tol", [("fp32", O.TINY, 2e-5), ("bf16", TOY64, 1e-2)])
def test_graph_replay_equals_eager_loop(prec, cfg, tol):
    m, _ = build_model(cfg, 3, prec)
    S = cfg["image_size"]
    y_cond, y_T, angle = _inputs(3, 4, S, 1)
    vc = torch.tensor([4, 2, 3])
    steps = list(range(1999, 1985, -1)) + [3, 2, 1, 0]            # two consecutive runs; t = 0 ends without noise
    zs = O.normal_draws(len(steps), (3, 3, S, S), seed=9)
    m.use_cuda_graph = False
    y_e, ret_e, *_ = m.generate(y_cond, vc, angle, y_t=y_T, noise_steps=zs, steps=steps)
    m.use_cuda_graph = True
    y_g, ret_g, *_ = m.generate(y_cond, vc, angle, y_t=y_T, noise_steps=zs, steps=steps)
    torch.cuda.synchronize()
    assert m._graph_error is None, m._graph_error
    plan = next(iter(m._plans.values()))
    assert len(plan.graphs) == 1, "a CUDA graph of the step must have been captured"
    assert margin(f"generate(): CUDA-graph replay vs eager loop, {prec}, 18 steps, injected noise: y rel-L2", rel(y_g, y_e), tol)
    assert ret_g.shape == ret_e.shape


@pytest.mark.parametrize("prec,cfg,tol", [("fp32", O.TINY, 2e-5), ("bf16", TOY64, 1e-2)])
def test_p_sample_graph_path_equals_eager_path(prec, cfg, tol):
    """p_sample(want_weights=False) replays the captured step (inputs copied into the plan's loop buffers, time-steps and seed
    into the device-resident state); with weights requested it runs eagerly.  Same injected noise -> same y_{t-1}, for ragged view
    counts, t > 0 and the noise-free t = 0 step; the caller's y_t must not be modified."""
    m, _ = build_model(cfg, 3, prec)
    S = cfg["image_size"]
    y_cond, y_T, angle = _inputs(3, 4, S, 4)
    vc = torch.tensor([4, 1, 3])
    z = torch.randn(3, 3, S, S, generator=torch.Generator().manual_seed(3)).cuda()
    for tval in (1500, 1, 0):
        t = torch.full((3,), tval, dtype=torch.long)
        keep = y_T.clone()
        y_e, logits, weights = m.p_sample(y_T, y_cond, vc, angle, t, noise=z)                      # eager (weights wanted); warms the plan
        assert logits is not None and weights is not None
        y_g, lg, wg = m.p_sample(y_T, y_cond, vc, angle, t, noise=z, want_weights=False)           # graph
        torch.cuda.synchronize()
        assert lg is None and wg is None and torch.equal(y_T, keep)
        assert margin(f"p_sample graph vs eager, {prec}, t={tval}: y_prev rel-L2", rel(y_g, y_e), tol)
    assert m._graph_error is None, m._graph_error
    plan = next(iter(m._plans.values()))
    assert len(plan.graphs) == 1, "p_sample must have captured / replayed a CUDA graph"
    # in-kernel noise: two calls draw different noise, re-seeding torch reproduces them
    t = torch.full((3,), 1000, dtype=torch.long)
    torch.manual_seed(11)
    a1 = m.p_sample(y_T, y_cond, vc, angle, t, want_weights=False)[0]
    a2 = m.p_sample(y_T, y_cond, vc, angle, t, want_weights=False)[0]
    torch.manual_seed(11)
    b1 = m.p_sample(y_T, y_cond, vc, angle, t, want_weights=False)[0]
    assert rel(a2, a1) > 1e-3 and rel(b1, a1) < 2e-5


def test_philox_noise_follows_torch_manual_seed():
    m, _ = build_model(O.TINY, 3, "fp32")
    y_cond, y_T, angle = _inputs(2, 3, 16, 2)
    vc = torch.tensor([3, 3])
    steps = list(range(1999, 1989, -1))
    outs = []
    for seed in (5, 5, 6):
        torch.manual_seed(seed)
        outs.append(m.generate(y_cond, vc, angle, y_t=y_T, steps=steps)[0])
    torch.cuda.synchronize()
    assert rel(outs[1], outs[0]) < 2e-5, "same torch seed -> same in-kernel noise (graph replays included)"
    assert rel(outs[2], outs[0]) > 1e-3, "another seed -> another noise stream"
    # and two instances seeded differently do not share a stream
    m2, _ = build_model(O.TINY, 3, "fp32")
    torch.manual_seed(7)
    a = m.generate(y_cond, vc, angle, y_t=y_T, steps=steps)[0]
    b = m2.generate(y_cond, vc, angle, y_t=y_T, steps=steps)[0]        # the generator advanced: a different seed is drawn
    assert rel(b, a) > 1e-3


def test_batch_shape_change_at_constant_image_count_fp32():
    """fp32 (CUDA-core) training: B = 2 with views [3, 1] and then B = 4 with views [1, 1, 1, 1] — the same 4 view-images, a
    different number of embedding rows.  Gradients of the second step must equal those of a fresh model."""
    cfg = O.TINY
    S = cfg["image_size"]

    def grads(m, B, vc, seed):
        g = torch.Generator().manual_seed(seed)
        y_cond = torch.rand(B, 3, 3, S, S, generator=g).cuda()
        y0 = torch.rand(B, 3, S, S, generator=g).cuda()
        noise = torch.randn(B, 3, S, S, generator=g).cuda()
        angle = torch.rand(B, 1, generator=g).cuda()
        t = torch.randint(1, 2000, (B,), generator=g)
        u = torch.rand(B, 1, generator=g)
        m.zero_grad(set_to_none=True)
        loss = m(y_cond=y_cond, view_count=torch.tensor(vc), angle=angle, y_0=y0, noise=noise, t=t, u=u)
        loss.backward()
        torch.cuda.synchronize()
        return float(loss.detach()), m.denoise_fn._flat_grad.clone()

    m1, _ = build_model(cfg, 4, "fp32")
    grads(m1, 2, [3, 1], 1)                       # leaves its activations / gradients in the arenas
    grads(m1, 3, [1, 1, 1], 2)                    # fewer images, other row count
    l_a, g_a = grads(m1, 4, [1, 1, 1, 1], 3)
    m2, _ = build_model(cfg, 4, "fp32")
    l_b, g_b = grads(m2, 4, [1, 1, 1, 1], 3)
    assert abs(l_a - l_b) < 1e-6 * abs(l_b)
    assert margin("fp32 training gradients after batch-shape changes in a reused arena vs fresh model: rel-L2", rel(g_a, g_b), 2e-5)


def test_backward_after_another_forward_is_refused():
    m, _ = build_model(O.TINY, 4, "fp32")
    S = 16
    g = torch.Generator().manual_seed(0)
    mk = lambda: dict(y_cond=torch.rand(2, 2, 3, S, S, generator=g).cuda(), view_count=torch.tensor([2, 1]), angle=torch.rand(2, 1, generator=g).cuda(),
                      y_0=torch.rand(2, 3, S, S, generator=g).cuda())
    l1 = m(**mk())
    l2 = m(**mk())
    with pytest.raises(RuntimeError, match="ANOTHER forward"):
        (l1 + l2).backward()
    m.zero_grad(set_to_none=True)
    l3 = m(**mk())
    l3.backward()                                  # the normal order still works
    torch.cuda.synchronize()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.denoise_fn.parameters())


def test_micro_batch_accumulation_equals_one_big_batch():
    """grad_accumulation(True): two half-batch backwards add natively into the same flat buffer == 2 x mean-of-halves."""
    m, _ = build_model(TOY64, 4, "bf16")
    S = 16
    g = torch.Generator().manual_seed(3)
    B = 4
    y_cond, y0 = torch.rand(B, 2, 3, S, S, generator=g).cuda(), torch.rand(B, 3, S, S, generator=g).cuda()
    noise, angle = torch.randn(B, 3, S, S, generator=g).cuda(), torch.rand(B, 1, generator=g).cuda()
    t, u = torch.randint(1, 2000, (B,), generator=g), torch.rand(B, 1, generator=g)
    vc = torch.full((B,), 2)

    def run(lo, hi):
        loss = m(y_cond=y_cond[lo:hi].contiguous(), view_count=vc[lo:hi], angle=angle[lo:hi].contiguous(), y_0=y0[lo:hi].contiguous(),
                 noise=noise[lo:hi].contiguous(), t=t[lo:hi], u=u[lo:hi])
        loss.backward()

    m.zero_grad(set_to_none=True)
    run(0, 2)
    g_a = m.denoise_fn._flat_grad.clone()
    m.zero_grad(set_to_none=True)
    run(2, 4)
    g_b = m.denoise_fn._flat_grad.clone()
    m.denoise_fn.grad_accumulation(True)
    m.zero_grad(set_to_none=True)
    run(0, 2)
    buf = m.denoise_fn._flat_grad
    run(2, 4)
    torch.cuda.synchronize()
    assert m.denoise_fn._flat_grad is buf, "the second backward must have added into the first one's buffer"
    p0 = next(m.denoise_fn.parameters())
    assert p0.grad.data_ptr() == buf.data_ptr()
    assert margin("micro-batch accumulation (bf16): accumulated gradient vs sum of separate backwards rel-L2", rel(buf, g_a + g_b), 1e-2)
    m.denoise_fn.grad_accumulation(False)


def test_phased_backward_equals_single_call():
    """vf_unet_backward_phase (the backward in n + 1 calls, for overlapping the data-parallel all-reduce) against vf_unet_backward:
    same gradients, phase-major flat layout, every phase's parameters contiguous."""
    m, _ = build_model(O.TINY, 4, "fp32")
    S = 16
    g = torch.Generator().manual_seed(1)
    B = 3
    kw = dict(y_cond=torch.rand(B, 3, 3, S, S, generator=g).cuda(), view_count=torch.tensor([3, 1, 2]), angle=torch.rand(B, 1, generator=g).cuda(),
              y_0=torch.rand(B, 3, S, S, generator=g).cuda(), noise=torch.randn(B, 3, S, S, generator=g).cuda(),
              t=torch.randint(1, 2000, (B,), generator=g), u=torch.rand(B, 1, generator=g))

    def run():
        m.zero_grad(set_to_none=True)
        m(**kw).backward()
        torch.cuda.synchronize()
        return {n: p.grad.detach().clone() for n, p in m.denoise_fn.named_parameters()}

    ref = run()
    unet = m.denoise_fn
    unet._grad_sync = (None, 4)            # 4 phases; no process group: the all-reduce calls are no-ops, the phasing is real
    unet._layout_cache = None
    got = run()
    order, bounds = unet._grad_layout(4)
    assert len(bounds) == 5 and bounds[-1][1] == unet._flat_grad.numel()
    assert all(b[1] > b[0] for b in bounds), "every phase owns parameters"
    # gradients that are analytically zero (the bias of a convolution feeding a GroupNorm whose groups hold ONE channel: C = 32,
    # 32 groups) come out as 1e-10 rounding noise: compare those on the scale of the largest gradient, the others relatively
    big = max(float(v.norm()) for v in ref.values())
    live = [n for n in ref if float(ref[n].norm()) > 1e-6 * big]
    assert all(float(got[n].norm()) < 1e-5 * big for n in ref if n not in live)
    worst = max(rel(got[n], ref[n]) for n in live)
    assert margin("phased backward (4 + 1 calls) vs single call, fp32: worst per-parameter rel-L2", worst, 2e-5)
    # the last phase holds exactly the embedding parameters
    names = [n for n, _ in unet.named_parameters()]
    plist = unet._params_in_order()
    pos = {id(p): i for i, p in enumerate(plist)}
    import ctypes as C
    from view_fusion_b200 import _lib
    ph = (C.c_int * len(plist))()
    _lib.check(_lib.load().vf_unet_backward_plan(unet._native(), 4, ph), "plan")
    late = {unet._param_names[i] for i in range(len(plist)) if ph[i] == 4}
    assert all(("noise_func" in n or n.startswith("noise_level_mlp")) for n in late) and len(late) == 4 + 2 * sum(1 for n in names if n.endswith("noise_func.noise_func.0.weight"))
    unet._grad_sync = None
    unet._layout_cache = None

"""Drop-in `ViewFusion` for the reference's `model/view_fusion.py`.

Keeps the constructor (:13-20), `set_new_noise_schedule` (:35-68), `forward` (:216-300), `generate` (:179-214),
`p_sample` (:166-177), `p_mean_variance` (:86-160), `q_sample`, `predict_start_from_noise`, `q_posterior`, the six
schedule buffers and the `denoise_fn.*` state_dict prefix.  The per-step work — view stacking, the shared UNet over
all views, softmax-over-views composition and the DDPM posterior update — is enqueued on the current CUDA stream
through the C ABI (`vf_pack_views` -> `vf_unet_forward` -> `vf_compose_ddpm_step`); the reference's host syncs
inside the step (`.tolist()`, tensor-valued `repeats`, `any(t > 0)`) are gone: `view_count` is read once per call.

Extra, optional keyword arguments (not in the reference) exist only to inject randomness for parity tests:
`p_sample(..., noise=)`, `generate(..., noise_steps=)`, `forward(..., t=, u=)`.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import _lib


def make_beta_schedule(schedule, num_timesteps, linear_start=1e-6, linear_end=1e-2, cosine_s=8e-3):
    """float64 beta schedules of view_fusion.py:330-362 (host-side numpy; runs once)."""
    lin = lambda a, b, n: np.linspace(a, b, n, dtype=np.float64)
    if schedule == "quad":
        return lin(linear_start ** 0.5, linear_end ** 0.5, num_timesteps) ** 2
    if schedule == "linear":
        return lin(linear_start, linear_end, num_timesteps)
    if schedule in ("warmup10", "warmup50"):
        betas = linear_end * np.ones(num_timesteps, dtype=np.float64)
        n = int(num_timesteps * (0.1 if schedule == "warmup10" else 0.5))
        betas[:n] = lin(linear_start, linear_end, n)
        return betas
    if schedule == "const":
        return linear_end * np.ones(num_timesteps, dtype=np.float64)
    if schedule == "jsd":
        return 1.0 / lin(num_timesteps, 1, num_timesteps)
    if schedule == "cosine":
        ts = np.arange(num_timesteps + 1, dtype=np.float64) / num_timesteps + cosine_s
        al = np.cos(ts / (1 + cosine_s) * math.pi / 2) ** 2
        al = al / al[0]
        return np.minimum(1 - al[1:] / al[:-1], 0.999)
    raise NotImplementedError(schedule)


class _Plan:
    """Per-(B, view_count) device state of the sampler / trainer: view offsets, staging buffers, and the device-resident
    loop state (`t_state`, noise counter, step record) that lets one captured CUDA graph serve every reverse step."""

    def __init__(self, model: "ViewFusion", y_cond: torch.Tensor, view_count: torch.Tensor):
        dev = y_cond.device
        vc = view_count.detach().to("cpu", torch.int64)            # the ONE host read of view_count per call
        self.B = int(vc.numel())
        self.vc = vc
        off = torch.zeros(self.B + 1, dtype=torch.int32)
        off[1:] = torch.cumsum(vc, 0).to(torch.int32)
        self.images = int(off[-1])
        self.max_v = int(vc.max())
        if int(vc.min()) < 1 or self.max_v > y_cond.shape[1]:
            raise ValueError("view_count must be within [1, y_cond.shape[1]]")
        if self.B != y_cond.shape[0]:
            raise ValueError("view_count must have one entry per sample")
        self.view_offset = off.to(dev)
        unet = model.denoise_fn
        S = unet.config["image_size"]
        self.H = self.W = S
        es = 2 if unet.precision == "bf16" else 4
        self.x0 = torch.empty(self.images * S * S * unet.k0 * es, dtype=torch.uint8, device=dev)
        self.img_sample = torch.empty(self.images, dtype=torch.int32, device=dev)
        self.out8 = torch.empty(self.images * S * S * 8, dtype=torch.float32, device=dev)
        self.t_state = torch.zeros(self.B, dtype=torch.int32, device=dev)
        self.level = torch.empty(self.B, dtype=torch.float32, device=dev)
        self.t_cur = torch.empty(self.B, dtype=torch.int32, device=dev)
        self.rec = torch.zeros(3, dtype=torch.int64, device=dev)           # vf_step_record (24 bytes)
        self.noise_ctr = torch.zeros(2, dtype=torch.int64, device=dev)     # {Philox offset, seed}: advanced on the device
        # reverse-loop state owned by the plan (generate): conditioning set, angles, y_t (updated in place), injected noise
        self.y_cond = self.angle = self.y = self.z = None
        self.graphs = {}
        self.warm = False

    def loop_buffers(self, y_cond, angle, y_t):
        if self.y_cond is None:
            self.y_cond, self.angle, self.y = torch.empty_like(y_cond), torch.empty_like(angle.reshape(-1)), torch.empty_like(y_t)
        self.y_cond.copy_(y_cond)
        self.angle.copy_(angle.reshape(-1))
        self.y.copy_(y_t)


class ViewFusion(nn.Module):
    def __init__(self, denoise_fn, beta_schedule, weighting_train=True, weighting_inference=True, **kwargs):
        super().__init__(**kwargs)
        self.denoise_fn = denoise_fn
        self.beta_schedule = beta_schedule
        self.loss_fn = F.mse_loss
        self.weighting_train = weighting_train
        self.weighting_inference = weighting_inference
        self.use_cuda_graph = True           # generate(): replay one captured graph per reverse step (eager when False)
        self._graph_error = None
        self._plans = {}
        print("Weighting train and inference:", self.weighting_train, self.weighting_inference)   # view_fusion.py:29-33

    # ---------------------------------------------------------------- schedule (view_fusion.py:35-68)
    def set_new_noise_schedule(self, device=torch.device("cuda"), phase="train"):
        betas = make_beta_schedule(**self.beta_schedule[phase])
        alphas = 1.0 - betas
        self.num_timesteps = int(betas.shape[0])
        gammas = np.cumprod(alphas, axis=0)
        gammas_prev = np.append(1.0, gammas[:-1])
        var = betas * (1.0 - gammas_prev) / (1.0 - gammas)
        tt = lambda a: torch.tensor(a, dtype=torch.float32, device=device)
        self.register_buffer("gammas", tt(gammas))
        self.register_buffer("sqrt_recip_gammas", tt(np.sqrt(1.0 / gammas)))
        self.register_buffer("sqrt_recipm1_gammas", tt(np.sqrt(1.0 / gammas - 1)))
        self.register_buffer("posterior_log_variance_clipped", tt(np.log(np.maximum(var, 1e-20))))
        self.register_buffer("posterior_mean_coef1", tt(betas * np.sqrt(gammas_prev) / (1.0 - gammas)))
        self.register_buffer("posterior_mean_coef2", tt((1.0 - gammas_prev) * np.sqrt(alphas) / (1.0 - gammas)))
        self._plans = {}                     # captured graphs hold the old buffers' addresses

    def _schedule_struct(self) -> _lib.Schedule:
        s = _lib.Schedule()
        for k in ("gammas", "sqrt_recip_gammas", "sqrt_recipm1_gammas", "posterior_log_variance_clipped",
                  "posterior_mean_coef1", "posterior_mean_coef2"):
            b = getattr(self, k)
            if not b.is_cuda:
                raise RuntimeError("schedule buffers must live on the GPU (set_new_noise_schedule(device='cuda'))")
            setattr(s, k, b.data_ptr())
        s.num_timesteps = self.num_timesteps
        return s

    # ---------------------------------------------------------------- small closed-form helpers (API parity)
    @staticmethod
    def _extract(a, t, ndim=4):
        return a.gather(-1, t).reshape(t.shape[0], *((1,) * (ndim - 1)))

    def predict_start_from_noise(self, y_t, t, noise):                                   # :70-74
        return self._extract(self.sqrt_recip_gammas, t) * y_t - self._extract(self.sqrt_recipm1_gammas, t) * noise

    def q_posterior(self, y_0_hat, y_t, t):                                              # :76-84
        mean = self._extract(self.posterior_mean_coef1, t) * y_0_hat + self._extract(self.posterior_mean_coef2, t) * y_t
        return mean, self._extract(self.posterior_log_variance_clipped, t)

    def q_sample(self, y_0, sample_gammas, noise=None):                                  # :162-164
        lib = _lib.require_device()
        noise = torch.randn_like(y_0) if noise is None else noise
        g = sample_gammas.reshape(-1).contiguous().float()
        y_0c, nz = y_0.contiguous().float(), noise.contiguous().float()
        out = torch.empty_like(y_0c)
        with torch.cuda.device(y_0c.device):
            _lib.check(lib.vf_q_sample(y_0c.data_ptr(), nz.data_ptr(), g.data_ptr(), y_0c.shape[0], y_0c[0].numel(),
                                       out.data_ptr(), _lib.stream_handle()), "vf_q_sample")
        return out

    # ---------------------------------------------------------------- the fused step
    @staticmethod
    def _fresh_seed() -> int:
        """64-bit Philox seed drawn from torch's default generator: follows torch.manual_seed (reproducible under re-seeding
        like the reference's randn_like), differs between calls, instances and — with per-rank seeds — ranks; no device sync."""
        return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())

    def step_overhead_launches(self) -> int:
        """Kernels a reverse step enqueues around the UNet: step record, view stacking (2), composition + DDPM update."""
        return 4

    def _enqueue_step(self, plan: _Plan, y_cond, angle, y_t, y_prev, *, z=None, add_noise=2, clip=True, weighting=True, advance=0,
                      seed=0, eps_out=None, weights_out=None, logits_out=None):
        """vf_p_sample_step: step record -> view stacking -> UNet over all views -> composition -> DDPM update, one enqueue,
        no host decision, no host sync (the time-steps live in plan.t_state on the device)."""
        lib = _lib.require_device()
        unet = self.denoise_fn
        B, n_max, Cc, H, W = y_cond.shape
        packed = unet.packed_weights()
        ws = unet.workspace(plan.images)
        unet.set_stash(False)                  # sampling keeps nothing for a backward
        unet._last_images = plan.images
        a = _lib.SampleStepArgs()
        a.packed, a.workspace, a.workspace_bytes = packed.data_ptr(), ws.data_ptr(), ws.numel()
        a.y_cond, a.B, a.n_max, a.cond_channels, a.H, a.W, a.images = y_cond.data_ptr(), B, n_max, Cc, H, W, plan.images
        a.view_offset, a.angle, a.y_t, a.y_prev = plan.view_offset.data_ptr(), angle.data_ptr(), y_t.data_ptr(), y_prev.data_ptr()
        a.t_state, a.advance, a.noise_ctr, a.seed, a.z = plan.t_state.data_ptr(), int(advance), plan.noise_ctr.data_ptr(), seed, _lib.ptr(z)
        a.add_noise, a.clip_denoised, a.weighting = int(add_noise), int(clip), int(weighting)
        a.x0, a.img_sample, a.unet_out = plan.x0.data_ptr(), plan.img_sample.data_ptr(), plan.out8.data_ptr()
        a.level, a.t_cur, a.rec = plan.level.data_ptr(), plan.t_cur.data_ptr(), plan.rec.data_ptr()
        a.eps_out, a.weights_out, a.logits_out, a.max_v = _lib.ptr(eps_out), _lib.ptr(weights_out), _lib.ptr(logits_out), plan.max_v
        a.sched = self._schedule_struct()
        _lib.check(lib.vf_p_sample_step(unet._native(), C.byref(a), _lib.stream_handle()), "vf_p_sample_step")
        unet._fwd_gen = unet.forward_generation()

    def _step(self, plan: _Plan, y_t, y_cond, angle, t, y_prev, *, z=None, add_noise=True, clip=True, weighting=True,
              eps_out=None, weights_out=None, logits_out=None):
        """One step at caller-supplied time-steps `t` (B,) (p_sample / p_mean_variance)."""
        plan.t_state.copy_(t)
        seed = 0
        if z is None and add_noise:
            seed = self._fresh_seed()
        self._enqueue_step(plan, y_cond, angle.reshape(-1), y_t, y_prev, z=z, add_noise=int(bool(add_noise)), clip=clip, weighting=weighting,
                           advance=0, seed=seed, eps_out=eps_out, weights_out=weights_out, logits_out=logits_out)

    @staticmethod
    def _prep(y_cond, angle, y_t=None):
        if not y_cond.is_cuda:
            raise RuntimeError("view_fusion_b200.ViewFusion needs CUDA tensors; there is no CPU fallback")
        y_cond = y_cond.contiguous().float()
        angle = angle.to(y_cond.device).contiguous().float()
        if angle.numel() != y_cond.shape[0]:
            raise ValueError("angle must be (B, 1)")
        if y_t is not None:
            if y_t.device != y_cond.device:
                raise RuntimeError("y_t and y_cond must live on the same device")
            y_t = y_t.contiguous().float()
        return y_cond, angle, y_t

    def _check_device(self, y_cond):
        p = next(self.denoise_fn.parameters())
        if p.device != y_cond.device:
            raise RuntimeError(f"model parameters live on {p.device}, inputs on {y_cond.device}")

    def _plan_for(self, y_cond, view_count) -> "_Plan":
        """Offsets, staging buffers, loop state and captured graphs per (view_count, shape): repeated p_sample / generate calls
        reuse them (stream order makes the reuse safe) instead of re-allocating, re-uploading and re-capturing."""
        vc = view_count.detach().to("cpu", torch.int64)
        key = (tuple(vc.tolist()), tuple(y_cond.shape), str(y_cond.device), self.denoise_fn.precision)
        plan = self._plans.get(key)
        if plan is None:
            if len(self._plans) >= 4:                      # a few live shapes (bench legs, ragged tails); drop the oldest
                self._plans.pop(next(iter(self._plans)))
            plan = self._plans[key] = _Plan(self, y_cond, vc)
        return plan

    def release_buffers(self) -> None:
        """Drop cached plans (staging buffers, captured graphs) and the UNet workspaces."""
        self._plans = {}
        self.denoise_fn.release_buffers()

    # ---------------------------------------------------------------- p_mean_variance / p_sample (:86-177)
    @torch.no_grad()
    def p_mean_variance(self, y_t, y_cond, view_count, angle, t, clip_denoised: bool):
        y_cond, angle, y_t = self._prep(y_cond, angle, y_t)
        self._check_device(y_cond)
        with torch.cuda.device(y_cond.device):
            plan = self._plan_for(y_cond, view_count)
            t = t.to(y_cond.device).long()
            mean = torch.empty_like(y_t)
            H, W = y_t.shape[-2:]
            w = self.weighting_inference
            weights = torch.empty(plan.B, plan.max_v, 3, H, W, device=y_t.device) if w else None
            logits = torch.empty(plan.images, 3, H, W, device=y_t.device) if w else None
            self._step(plan, y_t, y_cond, angle, t, mean, add_noise=False, clip=clip_denoised, weighting=w,
                       weights_out=weights, logits_out=logits)
        return mean, self._extract(self.posterior_log_variance_clipped, t), logits, weights

    @torch.no_grad()
    def p_sample(self, y_t, y_cond, view_count, angle, t, clip_denoised=True, noise=None, _plan=None, _eps_out=None,
                 want_weights=True):
        y_cond, angle, y_t = self._prep(y_cond, angle, y_t)
        self._check_device(y_cond)
        with torch.cuda.device(y_cond.device):
            plan = _plan if _plan is not None else self._plan_for(y_cond, view_count)
            # the reference's `any(t > 0)` (:176) is a host decision: a host-resident `t` answers it without touching the
            # device; a device-resident one costs the same sync the reference pays (generate() below decides on the device)
            add_noise = bool((t > 0).any())
            t = t.to(y_cond.device, non_blocking=True)
            H, W = y_t.shape[-2:]
            w = self.weighting_inference
            weights = torch.empty(plan.B, plan.max_v, 3, H, W, device=y_t.device) if (w and want_weights) else None
            logits = torch.empty(plan.images, 3, H, W, device=y_t.device) if (w and want_weights) else None
            if not add_noise:
                noise = None
            if self._p_sample_graph(plan, y_t, y_cond, angle, t, noise, clip_denoised, w, weights is not None or logits is not None
                                    or _eps_out is not None):
                return plan.y.clone(), None, None
            y_prev = torch.empty_like(y_t)
            self._step(plan, y_t, y_cond, angle, t, y_prev, z=None if noise is None else noise.to(y_t.device).contiguous().float(),
                       add_noise=add_noise, clip=clip_denoised, weighting=w, eps_out=_eps_out, weights_out=weights,
                       logits_out=logits)
            plan.warm = True
        return y_prev, logits, weights

    def _p_sample_graph(self, plan: _Plan, y_t, y_cond, angle, t, noise, clip, w, wants_extras: bool) -> bool:
        """p_sample through the captured step of generate(): the inputs are copied into the plan's loop buffers (device to
        device), the time-steps and the Philox seed into the device-resident loop state, and ONE graph replay runs the
        169 kernels; y_{t-1} is left in plan.y.  Taken when the caller wants nothing but y_{t-1} (no weights / logits /
        eps), clipping is on, and the plan has already run one eager step (lazy initialisations happen outside a capture).
        `any(t > 0)` is then decided on the device like in generate().  Returns False when the eager path must run."""
        unet = self.denoise_fn
        if wants_extras or not clip or not self.use_cuda_graph or unet._profiling or not plan.warm:
            return False
        with_z = noise is not None
        plan.loop_buffers(y_cond, angle, y_t)
        if with_z:
            if plan.z is None:
                plan.z = torch.empty_like(plan.y)
            plan.z.copy_(noise.to(y_t.device).float())
        plan.t_state.copy_(t)
        plan.noise_ctr.copy_(torch.tensor([0, 0 if with_z else self._fresh_seed()], dtype=torch.int64), non_blocking=True)
        gkey = (unet.packed_weights().data_ptr(), unet.workspace(plan.images).data_ptr(), with_z, w)
        return self._replay(plan, gkey, with_z, w)

    # ---------------------------------------------------------------- generate (:179-214)
    def _replay(self, plan: _Plan, key, with_z: bool, w: bool):
        """One reverse step through the plan's captured CUDA graph (captured on first use; any failure falls back to eager)."""
        g = plan.graphs.get(key)
        if g is None:
            plan.graphs.clear()                 # another workspace / weight buffer: the old captures are dead weight
            g = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(g):
                    self._enqueue_step(plan, plan.y_cond, plan.angle, plan.y, plan.y, z=plan.z if with_z else None, add_noise=2,
                                       clip=True, weighting=w, advance=1, seed=0)
            except Exception as e:              # capture unsupported on this driver: remember why, run eagerly from now on
                self._graph_error = repr(e)
                self.use_cuda_graph = False
                return False
            plan.graphs[key] = g
        g.replay()
        return True

    @torch.no_grad()
    def generate(self, y_cond, view_count, angle, y_t=None, sample_num=8, noise_steps=None, steps=None):
        """T-step reverse loop.  `noise_steps[j]` (optional) is the N(0,1) draw injected at the j-th executed step;
        `steps` (optional) restricts the loop to a list of time-steps (tests, benchmarks).

        The loop state (y_t, the time-steps, the noise counter) lives on the device and is advanced by the kernels themselves,
        so a step is one enqueue with no host decision; after the first (eager) step the remaining ones replay ONE captured
        CUDA graph (`use_cuda_graph`), except the `sample_num` snapshot steps, which also write the weights / logits."""
        y_cond, angle, y_t = self._prep(y_cond, angle, y_t)
        self._check_device(y_cond)
        b = y_cond.shape[0]
        assert self.num_timesteps > sample_num, "num_timesteps must greater than sample_num"
        sample_inter = self.num_timesteps // sample_num
        dev = y_cond.device
        with torch.cuda.device(dev):
            if y_t is None:
                y_t = torch.randn_like(y_cond[:, 0, :3, ...]).contiguous()
            plan = self._plan_for(y_cond, view_count)
            plan.loop_buffers(y_cond, angle, y_t)
            H, W = y_t.shape[-2:]
            w = self.weighting_inference
            ret_arr, weight_arr, logit_arr = [y_t], [], []
            order = list(reversed(range(self.num_timesteps))) if steps is None else [int(i) for i in steps]
            with_z = noise_steps is not None
            if with_z and plan.z is None:
                plan.z = torch.empty_like(plan.y)
            # Philox (offset, seed) of this call live on the device: a captured step holds no per-call constant
            plan.noise_ctr.copy_(torch.tensor([0, 0 if with_z else self._fresh_seed()], dtype=torch.int64))
            unet = self.denoise_fn
            graphs = self.use_cuda_graph and not unet._profiling
            gkey = (unet.packed_weights().data_ptr(), unet.workspace(plan.images).data_ptr(), with_z, w)
            prev = None
            for j, i in enumerate(order):
                if prev is None or i != max(prev - 1, 0) or prev == 0:
                    plan.t_state.fill_(i)               # (re)start of a consecutive run; inside one the kernels count down
                prev = i
                snap = i % sample_inter == 0
                if with_z:
                    plan.z.copy_(noise_steps[j])
                if snap or not graphs or not plan.warm or not self._replay(plan, gkey, with_z, w):
                    weights = torch.empty(plan.B, plan.max_v, 3, H, W, device=dev) if (w and snap) else None
                    logits = torch.empty(plan.images, 3, H, W, device=dev) if (w and snap) else None
                    self._enqueue_step(plan, plan.y_cond, plan.angle, plan.y, plan.y, z=plan.z if with_z else None, add_noise=2,
                                       clip=True, weighting=w, advance=1, seed=0, weights_out=weights, logits_out=logits)
                    plan.warm = True
                    if snap:
                        ret_arr.append(plan.y.clone())
                        logit_arr.append(logits)
                        weight_arr.append(weights)
            y_fin = plan.y.clone()
            ret_arr = torch.stack(ret_arr, dim=1)
            generated_samples = ret_arr[:, -1, ...]
            if w and logit_arr:
                logit_arr = torch.stack(logit_arr, dim=1)
                weight_arr = torch.stack(weight_arr, dim=1)
        return y_fin, ret_arr, logit_arr, weight_arr, generated_samples

    # ---------------------------------------------------------------- forward (:216-300)
    def forward(self, y_cond, view_count, angle, y_0=None, noise=None, generate=False, t=None, u=None):
        """generate=True -> generate().  Otherwise the training branch (view_fusion.py:229-300): scalar MSE loss whose
        backward() fills the gradients of every UNet parameter through the hand-written CUDA backward.
        `t` (B,) and `u` (B,1) optionally inject the reference's randint / rand draws (parity tests)."""
        if generate:
            return self.generate(y_cond, view_count, angle)
        y_cond, angle, y_0 = self._prep(y_cond, angle, y_0)
        dev = y_0.device
        b = y_0.shape[0]
        # the reference's draw order: randint (:231) -> rand (:234) -> randn_like (:239)
        if t is None:
            t = torch.randint(1, self.num_timesteps, (b,), device=dev)
        t = t.to(dev).long()
        if u is None:
            u = torch.rand((b, 1), device=dev)
        u = u.to(dev).float()
        g1 = self.gammas.gather(-1, t - 1).view(b, 1)
        g2 = self.gammas.gather(-1, t).view(b, 1)
        sample_gammas = ((g2 - g1) * u + g1).view(b).contiguous()
        if noise is None:
            noise = torch.randn_like(y_0)
        noise = noise.to(dev).contiguous().float()
        params = self.denoise_fn._params_in_order()
        return _TrainStep.apply(self, y_0, y_cond, view_count, angle, noise, sample_gammas, *params)


class _TrainStep(torch.autograd.Function):
    """loss = mse(noise, eps_hat(UNet over all views)); backward = hand-written CUDA backward of the whole step.

    The native plan keeps ONE tape / activation stash: the backward differentiates the most recent forward.  Each forward is
    stamped with the plan's forward generation and `backward` refuses to run against a stale stamp (e.g.
    `l1 = model(a); l2 = model(b); (l1 + l2).backward()`), instead of silently mixing forward #2's activations with forward
    #1's output gradient."""

    @staticmethod
    def forward(ctx, model, y_0, y_cond, view_count, angle, noise, sample_gammas, *params):
        lib = _lib.require_device()
        unet = model.denoise_fn
        with torch.cuda.device(y_0.device):
            st = _lib.stream_handle()
            B, n_max, Cc, H, W = y_cond.shape
            y_noisy = model.q_sample(y_0, sample_gammas, noise)                              # :240-242
            plan = _Plan(model, y_cond, view_count)
            _lib.check(lib.vf_pack_views(y_cond.data_ptr(), y_noisy.data_ptr(), plan.view_offset.data_ptr(), B, n_max, Cc, H, W,
                                         plan.images, unet.k0, unet.act_dtype, plan.x0.data_ptr(), plan.img_sample.data_ptr(), st),
                       "vf_pack_views")
            unet._last_images = plan.images
            ang = angle.reshape(-1).contiguous()
            unet.run_packed(plan.x0, plan.images, sample_gammas, ang, plan.img_sample, plan.out8, stash=True)   # grad mode is off in here
            loss = torch.zeros(1, dtype=torch.float32, device=y_0.device)
            grad8 = torch.empty_like(plan.out8)
            weighting = int(model.weighting_train)
            _lib.check(lib.vf_compose_mse(plan.out8.data_ptr(), plan.view_offset.data_ptr(), noise.data_ptr(), B, H, W, weighting,
                                          loss.data_ptr(), 0, grad8.data_ptr(), 1.0, st), "vf_compose_mse")
        ctx.model = model
        ctx.grad8 = grad8
        ctx.gen = unet._fwd_gen
        ctx.keep = (plan, ang, sample_gammas, y_noisy)          # device buffers the recorded tape points into
        ctx.n_params = len(params)
        return loss[0]

    @staticmethod
    def backward(ctx, grad_loss):
        unet = ctx.model.denoise_fn
        if unet.forward_generation() != ctx.gen:
            raise RuntimeError("view_fusion_b200: backward() of a ViewFusion loss after ANOTHER forward ran on the same UNet; the "
                               "native plan keeps one activation stash — call loss.backward() before the next forward "
                               "(one in-flight training forward per UNet)")
        with torch.cuda.device(ctx.grad8.device):
            g8 = ctx.grad8 * grad_loss            # device-side scale: no host read of the incoming gradient
            _, grads, accumulated = unet.run_backward(g8)
        if accumulated:                           # gradients were added into the live .grad views natively
            return (None,) * (7 + ctx.n_params)
        return (None,) * 7 + tuple(grads)

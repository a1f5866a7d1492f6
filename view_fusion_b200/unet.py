"""Drop-in `UNet` for the reference's `model/unet.py` (constructor :9-21, forward :114-138).

Same constructor arguments, same `forward(x, angle, time)` contract, same `state_dict()` keys/shapes (SURVEY.md
Appendix C) and — because the parameter-holding leaf modules are created in the reference's order with torch's
default initialisers — bit-identical random-init weights under the same `torch.manual_seed`.

The module tree below only HOLDS parameters.  The arithmetic runs in the sm_100a library behind the C ABI
(`include/viewfusion_b200.h`): `forward` packs the NCHW input, enqueues `vf_unet_forward` on the current CUDA
stream and converts the NHWC result back.  There is no PyTorch/CPU execution path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch
from torch import nn

from . import _lib


class Swish(nn.Module):
    """Placeholder keeping `noise_level_mlp` indices (0, 2) of unet.py:27-32; never executed."""


def _holder(**children) -> nn.Module:
    m = nn.Module()
    for k, v in children.items():
        m.add_module(k, v)
    return m


def _block(dim, dim_out, groups):
    # Sequential(GroupNorm, Swish, Identity, Conv2d) -> parameters at indices 0 and 3 (unet.py:210-215)
    return _holder(block=nn.Sequential(nn.GroupNorm(groups, dim), Swish(), nn.Identity(), nn.Conv2d(dim, dim_out, 3, padding=1)))


def _res_attn_block(dim, dim_out, emb_dim, groups, with_attn):
    # creation order == unet.py:232-238 then :254-256 so the RNG stream matches the reference's constructor
    noise_func = _holder(noise_func=nn.Sequential(nn.Linear(emb_dim, dim_out)))
    block1 = _block(dim, dim_out, groups)
    block2 = _block(dim_out, dim_out, groups)
    res_conv = nn.Conv2d(dim, dim_out, 1) if dim != dim_out else nn.Identity()
    m = _holder(res_block=_holder(noise_func=noise_func, block1=block1, block2=block2, res_conv=res_conv))
    if with_attn:
        m.add_module("attn", _holder(norm=nn.GroupNorm(groups, dim_out), qkv=nn.Conv2d(dim_out, dim_out * 3, 1, bias=False),
                                     out=nn.Conv2d(dim_out, dim_out, 1)))
    return m


class UNet(nn.Module):
    def __init__(
        self,
        in_channel: int = 6,
        out_channel: Optional[int] = 3,
        inner_channel: int = 32,
        norm_groups: int = 32,
        channel_mults: Sequence[int] = (1, 2, 4, 8, 8),
        attn_res: Sequence[int] = (8,),
        res_blocks: int = 3,
        dropout: float = 0,
        with_noise_level_emb: bool = True,
        image_size: int = 128,
        precision: str = "bf16",
    ):
        super().__init__()
        if dropout != 0:
            raise NotImplementedError("dropout != 0 is not supported (every reference config uses 0)")
        if not with_noise_level_emb:
            raise NotImplementedError("with_noise_level_emb=False is not supported (every reference config uses True)")
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' (tcgen05 tensor cores) or 'fp32' (CUDA cores, reference precision)")
        self.precision = precision
        self.config = dict(in_channel=in_channel, out_channel=out_channel if out_channel is not None else in_channel,
                           inner_channel=inner_channel, norm_groups=norm_groups, channel_mults=tuple(channel_mults),
                           attn_res=tuple(attn_res), res_blocks=res_blocks, image_size=image_size)

        # ---- parameter tree, mirroring unet.py:24-112 ----
        self.noise_level_mlp = nn.Sequential(nn.Linear(inner_channel, inner_channel * 4), Swish(),
                                             nn.Linear(inner_channel * 4, inner_channel))
        self.noise_level_angle_mlp = nn.Sequential()
        num_mults = len(channel_mults)
        pre, feat, res = inner_channel, [inner_channel], image_size
        downs = [nn.Conv2d(in_channel, inner_channel, kernel_size=3, padding=1)]
        for ind in range(num_mults):
            cm = inner_channel * channel_mults[ind]
            for _ in range(res_blocks):
                downs.append(_res_attn_block(pre, cm, inner_channel, norm_groups, res in attn_res))
                feat.append(cm)
                pre = cm
            if ind != num_mults - 1:
                downs.append(_holder(conv=nn.Conv2d(pre, pre, 3, 2, 1)))
                feat.append(pre)
                res //= 2
        self.downs = nn.ModuleList(downs)
        self.mid = nn.ModuleList([_res_attn_block(pre, pre, inner_channel, norm_groups, True),
                                  _res_attn_block(pre, pre, inner_channel, norm_groups, False)])
        ups = []
        for ind in reversed(range(num_mults)):
            cm = inner_channel * channel_mults[ind]
            for _ in range(res_blocks + 1):
                ups.append(_res_attn_block(pre + feat.pop(), cm, inner_channel, norm_groups, res in attn_res))
                pre = cm
            if ind >= 1:
                ups.append(_holder(conv=nn.Conv2d(pre, pre, 3, padding=1)))
                res *= 2
        self.ups = nn.ModuleList(ups)
        self.final_conv = _block(pre, self.config["out_channel"], norm_groups)

        # ---- native plan (host-side only; created lazily so CPU-side construction needs no GPU) ----
        self._plan = None
        self._packed = None
        self._packed_key = None
        self._packed_t = None
        self._packed_t_key = None
        self._ws = None
        self._capacity = 0            # view-images the workspaces are laid out for (grows, never shrinks)
        self._gws = None
        self._stash = True
        self._flat_grad = None
        self._plist = None
        self._fwd_gen = 0
        self._accumulate = False
        self._profiling = False
        self._flat_layout = 0

    # ------------------------------------------------------------------------------------------
    def _native(self):
        if self._plan is None:
            lib = _lib.require_device()
            cfg = _lib.UnetConfig()
            c = self.config
            cfg.in_channel, cfg.out_channel, cfg.inner_channel = c["in_channel"], c["out_channel"], c["inner_channel"]
            cfg.norm_groups, cfg.res_blocks, cfg.image_size = c["norm_groups"], c["res_blocks"], c["image_size"]
            cfg.n_mults = len(c["channel_mults"])
            for i, m in enumerate(c["channel_mults"]):
                cfg.channel_mults[i] = m
            cfg.n_attn_res = len(c["attn_res"])
            for i, r in enumerate(c["attn_res"]):
                cfg.attn_res[i] = r
            h = C.c_void_p()
            _lib.check(lib.vf_unet_create(C.byref(cfg), _lib.VF_BF16 if self.precision == "bf16" else _lib.VF_F32, C.byref(h)),
                       "vf_unet_create")
            self._plan = h
            # parameter table of the plan must be exactly this module's state_dict
            names = []
            buf = C.create_string_buffer(256)
            shape = (C.c_int64 * 4)()
            nd = C.c_int()
            mine = dict(self.named_parameters())
            for i in range(lib.vf_unet_num_params(h)):
                _lib.check(lib.vf_unet_param_info(h, i, buf, 256, shape, C.byref(nd)), "vf_unet_param_info")
                n = buf.value.decode()
                if n not in mine or tuple(mine[n].shape) != tuple(shape[: nd.value]):
                    raise RuntimeError(f"plan/module parameter mismatch at {n}")
                names.append(n)
            if len(names) != len(mine):
                raise RuntimeError("plan/module parameter count mismatch")
            self._param_names = names
            self._plist = [mine[n] for n in names]
        return self._plan

    def __del__(self):
        try:
            if getattr(self, "_plan", None) is not None:
                _lib.load().vf_unet_destroy(self._plan)
        except Exception:
            pass

    @property
    def act_dtype(self) -> int:
        return _lib.VF_BF16 if self.precision == "bf16" else _lib.VF_F32

    def packed_weights(self) -> torch.Tensor:
        """GEMM-ready weight cache; refreshed whenever a master parameter changed (optimizer step, load_state_dict)."""
        lib = _lib.require_device()
        h = self._native()
        plist = self._params_in_order()
        key = tuple((p.data_ptr(), p._version) for p in plist)
        if self._packed is None or key != self._packed_key:
            for p in plist:
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("view_fusion_b200.UNet parameters must be contiguous fp32 CUDA tensors (call .cuda())")
            dev = plist[0].device
            if self._packed is None or self._packed.device != dev:
                self._packed = torch.empty(lib.vf_unet_packed_bytes(h), dtype=torch.uint8, device=dev)
            arr = (C.c_void_p * len(plist))(*[p.data_ptr() for p in plist])
            _lib.check(lib.vf_unet_pack_weights(h, arr, self._packed.data_ptr(), _lib.stream_handle()), "vf_unet_pack_weights")
            self._packed_key = key
        return self._packed

    def workspace(self, images: int) -> torch.Tensor:
        """Zero-filled arena of the forward.  Padding rows of convolution outputs are never written and must read as zeros
        (3x3 halos, weight-gradient reductions), so the LAYOUT must not move between calls: the library sizes every buffer for
        the capacity (`vf_unet_set_capacity`), not for the current image count — a different batch size or different view
        counts (the reference draws view_count per step, experiment.py:277) reuse the same zeroed arena without a memset.  A
        larger request grows the capacity and allocates a fresh zero-filled arena (and drops the backward's)."""
        lib = _lib.require_device()
        h = self._native()
        dev = self._params_in_order()[0].device
        if self._ws is None or images > self._capacity or self._ws.device != dev:
            cap = max(self._capacity, images)
            _lib.check(lib.vf_unet_set_capacity(h, cap), "vf_unet_set_capacity")
            self._ws = self._gws = None               # release before allocating the larger arenas
            self._ws = torch.zeros(lib.vf_unet_workspace_bytes(h, cap), dtype=torch.uint8, device=dev)
            self._capacity = cap
        return self._ws

    def grad_accumulation(self, on: bool) -> None:
        """Micro-batching: with `on`, a backward whose parameters still hold the previous backward's flat-buffer gradients adds
        into that buffer natively (one launch chain, no 400 per-tensor adds); `optimizer.zero_grad(set_to_none=True)` starts over."""
        self._accumulate = bool(on)

    def release_buffers(self) -> None:
        """Drop the workspaces (they are re-created, zero-filled, on the next call)."""
        self._ws = self._gws = None
        self._capacity = 0
        if self._plan is not None:
            _lib.check(_lib.load().vf_unet_set_capacity(self._plan, 0), "vf_unet_set_capacity")

    def invalidate_packed(self) -> None:
        """Force a re-pack of the GEMM-ready weights on the next forward (after writes that bypass the version counter,
        e.g. `.data` updates)."""
        self._packed_key = None
        self._packed_t_key = None

    # ---------------------------------------------------------------- training support
    def _params_in_order(self):
        self._native()
        pl = self._plist
        # .to(device) / load_state_dict keep the Parameter objects; a re-registered parameter (assign=True, module surgery)
        # is caught by checking the two ends of the table
        if pl is None or pl[0] is not self._param_by_name(self._param_names[0]) or pl[-1] is not self._param_by_name(self._param_names[-1]):
            mine = dict(self.named_parameters())
            pl = self._plist = [mine[n] for n in self._param_names]
        return pl

    def _param_by_name(self, name: str):
        m = self
        parts = name.split(".")
        for q in parts[:-1]:
            m = m._modules[q]
        return m._parameters[parts[-1]]

    def packed_weights_t(self) -> torch.Tensor:
        """Transposed packs for the data gradients; refreshed together with the forward packs."""
        lib = _lib.require_device()
        h = self._native()
        self.packed_weights()
        if self._packed_t is None or self._packed_t_key != self._packed_key:
            dev = self._packed.device
            if self._packed_t is None or self._packed_t.device != dev:
                self._packed_t = torch.empty(lib.vf_unet_packed_t_bytes(h), dtype=torch.uint8, device=dev)
            _lib.check(lib.vf_unet_pack_weights_t(h, self._packed_t.data_ptr(), _lib.stream_handle()), "vf_unet_pack_weights_t")
            self._packed_t_key = self._packed_key
        return self._packed_t

    def run_backward(self, grad_out8: torch.Tensor):
        """Backward of the last run_packed: returns per-parameter fp32 gradients (views of one flat buffer)."""
        lib = _lib.require_device()
        h = self._native()
        plist = self._params_in_order()
        pt = self.packed_weights_t()
        dev = plist[0].device
        nbytes = lib.vf_unet_backward_workspace_bytes(h)          # laid out for the capacity, like the forward arena
        if self._gws is None or self._gws.numel() < nbytes or self._gws.device != dev:
            self._gws = None
            self._gws = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        # every gradient starts on a 16-byte boundary of the flat buffer (vectorised optimizer / all-reduce accesses)
        sync = getattr(self, "_grad_sync", None)
        syncing = sync is not None and not getattr(self, "_no_sync", False)
        n_ph = int(sync[1]) if sync is not None else 0
        order, bounds = self._grad_layout(n_ph)       # parameter order inside the flat buffer (+ slice of every phase)
        # micro-batch accumulation (grad_accumulation(True)): when every .grad still IS its slice of the previous flat buffer,
        # the library adds into that buffer and autograd gets nothing to add a second time
        accumulate = False
        if self._accumulate and self._flat_grad is not None and self._flat_grad.device == dev and self._flat_layout == n_ph:
            g0 = plist[order[0]].grad
            accumulate = g0 is not None and g0.data_ptr() == self._flat_grad.data_ptr()
        total = bounds[-1][1] if bounds else sum((p.numel() + 3) & ~3 for p in plist)
        flat = self._flat_grad if accumulate else torch.zeros(total, dtype=torch.float32, device=dev)
        grads, off = [None] * len(plist), 0
        for i in order:
            p = plist[i]
            grads[i] = flat[off:off + p.numel()].view_as(p)
            off += (p.numel() + 3) & ~3
        arr = (C.c_void_p * len(grads))(*[g.data_ptr() for g in grads])
        st = _lib.stream_handle()
        if n_ph == 0:
            _lib.check(lib.vf_unet_backward(h, pt.data_ptr(), self._gws.data_ptr(), self._gws.numel(), grad_out8.data_ptr(), arr, st),
                       "vf_unet_backward")
        else:
            # data parallel: the backward runs in n_ph + 1 phases; the all-reduce of a phase's slice of the flat gradient is issued
            # the moment that phase is enqueued, so NCCL (its own stream, ordered after the work enqueued so far) exchanges the
            # decoder-side gradients over NVLink while the encoder-side layers are still being differentiated
            from .distributed import allreduce_mean_async
            works = []
            for k in range(n_ph + 1):
                _lib.check(lib.vf_unet_backward_phase(h, pt.data_ptr(), self._gws.data_ptr(), self._gws.numel(), grad_out8.data_ptr(), arr,
                                                      k, n_ph, st), "vf_unet_backward_phase")
                if syncing and bounds[k][1] > bounds[k][0]:
                    works.append(allreduce_mean_async(flat[bounds[k][0]:bounds[k][1]], sync[0]))
            for w in works:
                w()
        self._flat_grad = flat
        self._flat_layout = n_ph
        return plist, grads, accumulate

    def _grad_layout(self, n_ph: int):
        """Order of the parameters inside the flat gradient buffer and the [start, end) float range of every backward phase.
        n_ph == 0: parameter order, one range.  Otherwise phase-major (vf_unet_backward_plan: phase k = k-th run of the reversed
        tape, phase n_ph = embedding parameters), so that each phase's gradients are ONE contiguous all-reduce."""
        cached = getattr(self, "_layout_cache", None)
        if cached is not None and cached[0] == n_ph:
            return cached[1], cached[2]
        plist = self._params_in_order()
        if n_ph == 0:
            order, bounds = list(range(len(plist))), []
        else:
            ph = (C.c_int * len(plist))()
            _lib.check(_lib.load().vf_unet_backward_plan(self._native(), n_ph, ph), "vf_unet_backward_plan")
            order = sorted(range(len(plist)), key=lambda i: (ph[i], i))
            bounds, off = [], 0
            for k in range(n_ph + 1):
                start = off
                for i in order:
                    if ph[i] == k:
                        off += (plist[i].numel() + 3) & ~3
                bounds.append((start, off))
        self._layout_cache = (n_ph, order, bounds)
        return order, bounds

    @property
    def k0(self) -> int:
        return _lib.require_device().vf_unet_k0(self._native())

    def run_packed(self, x0: torch.Tensor, images: int, level: torch.Tensor, angle: torch.Tensor, img_row: torch.Tensor,
                   out: torch.Tensor, stash=None) -> None:
        """Enqueue one UNet forward on pre-packed input (the sampler / trainer entry)."""
        lib = _lib.require_device()
        h = self._native()
        packed = self.packed_weights()
        ws = self.workspace(images)
        if stash is None:
            stash = torch.is_grad_enabled()          # under no_grad nothing extra is kept for vf_unet_backward
        stash = bool(stash)
        if stash != self._stash:
            _lib.check(lib.vf_unet_set_stash(h, int(stash)), "vf_unet_set_stash")
            self._stash = stash
        _lib.check(lib.vf_unet_forward(h, packed.data_ptr(), ws.data_ptr(), ws.numel(), images, x0.data_ptr(), level.data_ptr(),
                                       angle.data_ptr(), level.numel(), img_row.data_ptr(), out.data_ptr(), _lib.stream_handle()),
                   "vf_unet_forward")
        self._fwd_gen = int(lib.vf_unet_forward_generation(h))

    def forward_generation(self) -> int:
        """Counter of forwards run on the native plan: the backward differentiates the LAST one."""
        return int(_lib.load().vf_unet_forward_generation(self._native()))

    def set_stash(self, stash: bool) -> None:
        if bool(stash) != self._stash:
            _lib.check(_lib.load().vf_unet_set_stash(self._native(), int(bool(stash))), "vf_unet_set_stash")
            self._stash = bool(stash)

    def last_launches(self) -> int:
        return _lib.load().vf_unet_last_launches(self._native())

    KERNEL_CLASSES = ("conv", "gn_stats", "gn_apply", "attention", "upsample", "embed")

    def set_profiling(self, on: bool) -> None:
        _lib.check(_lib.load().vf_unet_set_profiling(self._native(), int(on)), "vf_unet_set_profiling")
        self._profiling = bool(on)

    def profile(self) -> dict:
        """{class: (ms, launches)} of the last forward run with profiling on (synchronises)."""
        ms = (C.c_float * 6)()
        cnt = (C.c_int * 6)()
        _lib.check(_lib.load().vf_unet_profile_read(self._native(), ms, cnt), "vf_unet_profile_read")
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(self.KERNEL_CLASSES)}

    def profile_launches(self, cap: int = 1024) -> list:
        """[(class, ms, images, H, C_in_or_K, C_out, ksize, stride)] per launch of the last profiled forward."""
        ms = (C.c_float * cap)()
        kind = (C.c_int * cap)()
        desc = (C.c_int * (6 * cap))()
        n = _lib.load().vf_unet_profile_launches(self._native(), ms, kind, desc, cap)
        if n < 0:
            _lib.check(n, "vf_unet_profile_launches")
        return [(self.KERNEL_CLASSES[kind[j]] if 0 <= kind[j] < len(self.KERNEL_CLASSES) else "other", float(ms[j]),
                 *[int(desc[6 * j + q]) for q in range(6)]) for j in range(n)]

    def read_tap(self, name: str) -> torch.Tensor:
        """Output activation of module `name` in the last forward, as NCHW fp32 (parity debugging)."""
        lib = _lib.require_device()
        h = self._native()
        chw = (C.c_int64 * 3)()
        # size query: taps are at most images * Cmax * S * S; allocate generously then trim
        S = self.config["image_size"]
        cmax = 2 * self.config["inner_channel"] * max(self.config["channel_mults"])
        buf = torch.empty(self._last_images * cmax * S * S, dtype=torch.float32, device=self._ws.device)
        # taps are addressed inside the CURRENT workspace: the last forward must have used it
        _lib.check(lib.vf_unet_read_tap(h, self._ws.data_ptr(), name.encode(), buf.data_ptr(), chw, _lib.stream_handle()), "vf_unet_read_tap")
        c, hh, ww = chw[0], chw[1], chw[2]
        return buf[: self._last_images * c * hh * ww].view(self._last_images, c, hh, ww).clone()

    # ------------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, angle: torch.Tensor, time: torch.Tensor) -> torch.Tensor:
        """x (R, Cin, H, W) fp32, angle (R, 1), time (R, 1) [the noise level] -> (R, Cout, H, W) fp32 (unet.py:114-138)."""
        lib = _lib.require_device()
        if not x.is_cuda:
            raise RuntimeError("view_fusion_b200.UNet.forward needs CUDA tensors; there is no CPU fallback")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and x.requires_grad:
            raise NotImplementedError("gradients w.r.t. the UNet input are not provided")
        R, Cin, H, W = x.shape
        S = self.config["image_size"]
        if (H, W) != (S, S) or Cin != self.config["in_channel"]:
            raise ValueError(f"expected input (R,{self.config['in_channel']},{S},{S}), got {tuple(x.shape)}")
        if self._params_in_order()[0].device != x.device:
            raise RuntimeError(f"UNet parameters live on {self._params_in_order()[0].device}, the input on {x.device}")
        x = x.contiguous().float()
        with torch.cuda.device(x.device):          # launches follow the tensors' device, not the caller's current one
            k0 = self.k0
            es = 2 if self.precision == "bf16" else 4
            x0 = torch.empty(R * H * W * k0 * es, dtype=torch.uint8, device=x.device)
            st = _lib.stream_handle()
            _lib.check(lib.vf_pack_nchw(x.data_ptr(), R, Cin, H, W, k0, self.act_dtype, x0.data_ptr(), st), "vf_pack_nchw")
            level = time.to(x.device).reshape(-1).contiguous().float()
            ang = angle.to(x.device).reshape(-1).contiguous().float()
            if level.numel() != R or ang.numel() != R:
                raise ValueError("angle and time must be (R, 1)")
            img_row = torch.arange(R, dtype=torch.int32, device=x.device)
            out8 = torch.empty(R * H * W * 8, dtype=torch.float32, device=x.device)
            self._last_images = R
            self.run_packed(x0, R, level, ang, img_row, out8)
            oc = self.config["out_channel"]
            out = torch.empty(R, oc, H, W, dtype=torch.float32, device=x.device)
            _lib.check(lib.vf_nhwc_to_nchw(out8.data_ptr(), 8, R, oc, H, W, out.data_ptr(), st), "vf_nhwc_to_nchw")
        return out

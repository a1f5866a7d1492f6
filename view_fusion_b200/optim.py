"""Fused Adam for the ViewFusion training step (reference: `torch.optim.Adam` built in
experiment.py:115-120 and stepped at :293, learning rate rewritten every iteration by utils/schedulers.py:10-14 through
`param_groups[i]["lr"]`).

Same constructor defaults, update rule, `state` layout (`step`, `exp_avg`, `exp_avg_sq` per parameter) and
`state_dict()` as `torch.optim.Adam`, so the reference's checkpoint code (`utils/checkpoint.py`) keeps working; the step
itself is ONE CUDA launch over all ~400 parameter tensors (`vf_adam_step`) instead of torch's ~60 multi-tensor launches.
No amsgrad / maximize / capturable variants (the reference does not use them)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class _Entry(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("numel", C.c_int), ("reserved", C.c_int)]


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0.0 or eps < 0.0 or not (0.0 <= betas[0] < 1.0) or not (0.0 <= betas[1] < 1.0) or weight_decay < 0.0:
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._tables = {}

    def _table(self, gi, ps):
        key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr(), self.state[p]["exp_avg_sq"].data_ptr(),
                     p.numel()) for p in ps)
        # gradients usually alternate between a few flat buffers (the caching allocator hands back the same blocks), so a
        # handful of tables per group covers every later step without rebuilding or re-uploading anything
        per_group = self._tables.setdefault(gi, {})
        cached = per_group.get(key)
        if cached is not None:
            return cached
        lib = _lib.require_device()
        chunk = lib.vf_adam_chunk_elems()
        entries = (_Entry * len(ps))()
        chunk_entry, chunk_start = [], []
        for i, (pp, gp, mp, vp, n) in enumerate(key):
            entries[i] = _Entry(pp, gp, mp, vp, n, 0)
            for s in range(0, n, chunk):
                chunk_entry.append(i)
                chunk_start.append(s)
        dev = ps[0].device
        raw = torch.frombuffer(bytearray(bytes(entries)), dtype=torch.uint8).to(dev)
        ce = torch.tensor(chunk_entry, dtype=torch.int32, device=dev)
        cs = torch.tensor(chunk_start, dtype=torch.int32, device=dev)
        if len(per_group) >= 8:
            per_group.pop(next(iter(per_group)))
        per_group[key] = (raw, ce, cs)
        return raw, ce, cs

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.require_device()
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous() or p.grad.dtype != torch.float32:
                    raise RuntimeError("FusedAdam needs contiguous fp32 CUDA parameters and gradients")
                if p.grad.is_sparse:
                    raise RuntimeError("FusedAdam does not support sparse gradients")
                if not p.grad.is_contiguous():
                    p.grad = p.grad.contiguous()
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            steps = {int(self.state[p]["step"]) for p in ps}
            if len(steps) != 1:
                raise RuntimeError("FusedAdam: parameters of one group must share the step count")
            step = steps.pop() + 1
            raw, ce, cs = self._table(gi, ps)
            b1, b2 = group["betas"]
            _lib.check(lib.vf_adam_step(raw.data_ptr(), ce.data_ptr(), cs.data_ptr(), int(ce.numel()), float(group["lr"]), float(b1),
                                        float(b2), float(group["eps"]), float(group["weight_decay"]), step, _lib.stream_handle()),
                       "vf_adam_step")
            for p in ps:
                self.state[p]["step"] += 1
            # the kernel wrote through raw pointers: tell autograd / the packed-weight cache that the parameters changed
            torch.autograd.graph.increment_version(ps)
        return loss


class LrScheduler:
    """Learning-rate schedule of the reference's training loop (utils/schedulers.py:1-14, used at experiment.py:115-120 and
    :265-267): linear warm-up to `peak_lr` over `peak_it` iterations, then `peak_lr * decay_rate ** ((it - peak_it) / decay_it)`.
    Same constructor, same `get_cur_lr(it)`; `apply(optimizer, it)` is the three-line rewrite of `param_groups[i]["lr"]` the
    reference's loop does every iteration (FusedAdam reads the value at the next step)."""

    def __init__(self, peak_lr=4e-4, peak_it=10000, decay_rate=0.5, decay_it=100000):
        self.peak_lr = peak_lr
        self.peak_it = peak_it
        self.decay_rate = decay_rate
        self.decay_it = decay_it

    def get_cur_lr(self, it):
        if it < self.peak_it:
            return self.peak_lr * (it / self.peak_it)
        return self.peak_lr * (self.decay_rate ** ((it - self.peak_it) / self.decay_it))

    def apply(self, optimizer, it):
        lr = self.get_cur_lr(it)
        for group in optimizer.param_groups:
            group["lr"] = lr
        return lr

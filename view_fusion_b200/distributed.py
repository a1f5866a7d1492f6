"""Data-parallel plumbing for the ViewFusion hot path (reference: experiment.py:104-110 DDP wrap, :159-160 per-rank
batch, utils/dist.py:11-26 process-group init).

The path shards by SAMPLE: all views of a sample stay on one GPU (they meet in the softmax-over-views composition),
samples never interact.  Sampling therefore needs no collective at all; training needs exactly one exchange per step,
the mean of the parameter gradients.  The UNet backward writes every parameter gradient into ONE flat fp32 buffer
(`UNet.run_backward`), so the exchange is a handful of large `all_reduce` calls on slices of that buffer instead of
DDP's bucket copies; slices are issued tail-first because the backward finishes the decoder-side (tail) gradients
first.  The same code runs on `gloo` CPU tensors (tests) and on NCCL over NVLink (B200).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_samples(view_count: Sequence[int], world_size: int) -> List[Tuple[int, int]]:
    """Contiguous [start, end) sample ranges per rank, balanced by the number of view-images (sum of view_count).

    Every rank gets at least one sample when len(view_count) >= world_size; with equal view counts this is the
    reference's `batch_size // world_size` split (experiment.py:159-160)."""
    n = len(view_count)
    if world_size <= 0:
        raise ValueError("world_size must be positive")
    if n < world_size:
        raise ValueError(f"cannot shard {n} samples over {world_size} ranks")
    vc = [int(v) for v in view_count]
    total = sum(vc)
    bounds, start, acc = [], 0, 0
    for r in range(world_size):
        remaining_ranks = world_size - r
        if remaining_ranks == 1:
            end = n
        else:
            target = (total - acc) / remaining_ranks
            end, run = start, 0
            # take samples while that brings this rank closer to its share, leaving >= 1 sample per later rank
            while end < n - (remaining_ranks - 1):
                nxt = run + vc[end]
                if end > start and abs(nxt - target) > abs(run - target):
                    break
                run = nxt
                end += 1
            end = max(end, start + 1)
        bounds.append((start, end))
        acc += sum(vc[start:end])
        start = end
    return bounds


def chunk_bounds(numel: int, chunks: int, align: int = 1024) -> List[Tuple[int, int]]:
    """Split [0, numel) into <= `chunks` aligned slices (the last one takes the remainder)."""
    chunks = max(1, int(chunks))
    per = -(-numel // chunks)
    per = -(-per // align) * align
    out, s = [], 0
    while s < numel:
        e = min(numel, s + per)
        out.append((s, e))
        s = e
    return out


def allreduce_mean_(flat: torch.Tensor, group: Optional[dist.ProcessGroup] = None, chunks: int = 4) -> torch.Tensor:
    """In-place mean over the ranks of a flat gradient buffer, issued as `chunks` slices, tail first."""
    if not dist.is_available() or not dist.is_initialized():
        return flat
    world = dist.get_world_size(group)
    if world == 1:
        return flat
    waits = [allreduce_mean_async(flat[s:e], group) for s, e in reversed(chunk_bounds(flat.numel(), chunks))]
    for w in waits:
        w()
    return flat


def allreduce_mean_async(t: torch.Tensor, group: Optional[dist.ProcessGroup] = None):
    """Start the in-place mean of `t` over the ranks; returns a callable that makes the current stream (or the host, for CPU
    backends) wait for it.  NCCL averages in the collective itself (ReduceOp.AVG: no separate scaling pass); the collective is
    ordered after everything already enqueued on the current stream and runs on NCCL's own stream."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return lambda: None
    world = dist.get_world_size(group)
    if t.is_cuda and dist.get_backend(group) == "nccl":
        w = dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group, async_op=True)
        return w.wait
    w = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=True)

    def finish():
        w.wait()
        t.mul_(1.0 / world)
    return finish


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group: Optional[dist.ProcessGroup] = None) -> None:
    """Start-up parameter broadcast (what the DDP constructor does, experiment.py:105-107)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    with torch.no_grad():
        for p in module.parameters():
            dist.broadcast(p.detach(), src=src, group=group)     # shares p's version counter (p.data would not)
    for m in module.modules():
        if hasattr(m, "invalidate_packed"):
            m.invalidate_packed()                                  # derived bf16 weight packs follow the new masters


def data_parallel(model, group: Optional[dist.ProcessGroup] = None, chunks: int = 4):
    """Turn on the per-step gradient all-reduce for a view_fusion_b200.ViewFusion (or UNet) and sync its parameters.

    Replaces `DistributedDataParallel(model, device_ids=[rank])` of the reference: call once after `.to(device)`;
    `loss.backward()` then leaves rank-averaged gradients in `param.grad`."""
    unet = getattr(model, "denoise_fn", model)
    broadcast_parameters(unet, 0, group)
    unet._grad_sync = (group, int(chunks))        # chunks = backward phases whose all-reduce overlaps the rest of the backward
    unet._layout_cache = None
    return model


class no_sync:
    """Context manager: backwards inside it skip the gradient exchange (micro-batch accumulation; DDP.no_sync of the reference's
    wrapper).  The last micro-batch runs outside and exchanges the accumulated gradients."""

    def __init__(self, model):
        self.unet = getattr(model, "denoise_fn", model)

    def __enter__(self):
        self.unet._no_sync = True
        return self

    def __exit__(self, *a):
        self.unet._no_sync = False
        return False

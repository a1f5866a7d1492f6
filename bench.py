#!/usr/bin/env python
"""Headline benchmark of the ViewFusion hot path on B200 (BASELINE.json metric: sampled target views/sec).

Workload (configs/small-v100.yaml of the reference = BASELINE.json configs[1]): small UNet, 64x64, B=28 samples
per GPU, N=6 conditioning views (168 view-images through the shared UNet per step), T=2000 reverse steps.

  step   = ONE reverse-diffusion step over the batch: view stacking -> UNet over all views -> softmax-over-views
           composition -> DDPM posterior update (ViewFusion.p_sample's work).  Every one of the T steps of
           `generate` is this same work, so   views/s = n_gpus * B / (T * seconds_per_step).
  value  = device-timed (CUDA events), inputs resident in HBM: K consecutive steps of the reverse loop run through the
           public `ViewFusion.generate(..., steps=[T-1, T-2, ...])`; K = T is literally one full generate().
  generate_full = ONE real `model(y_cond=, view_count=, angle=, generate=True)` call (all T = 2000 steps, the call
           experiment.py:337-342 makes), timed on the device and by wall clock, next to the K-step extrapolation.
  library_baseline = the same step run by the library kernels torch dispatches on this GPU (cuDNN / cuBLAS / ATen through
           the functional restatement of the reference graph): fp32 with TF32 off, and bf16 autocast + channels_last.
  e2e    = the same metric through the public API `ViewFusion.p_sample` with HOST buffers: every step copies
           y_cond / y_t / angle / view_count from pinned host memory and reads y_{t-1} back.
  roofline      = the dominant kernel class (tcgen05 implicit-GEMM convolution): algorithmic conv FLOPs per
           step / summed CUDA-event durations of its launches in one profiled step, against MEASURED_PEAKS.json.
  cpu_baseline  = the oracle (CPU fp32 restatement of the reference, oracle/vf_oracle.py) on the host cores.
  --impl reference : times that CPU implementation alone, on a bounded sample of the same workload.

One process per GPU; sampling shards by sample batch with no collective (weak scaling).
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

SMALL = dict(in_channel=6, out_channel=6, inner_channel=64, norm_groups=32, channel_mults=(1, 2, 3, 5),
             attn_res=(16,), res_blocks=3, image_size=64)
BETA = {"train": dict(schedule="linear", num_timesteps=2000, linear_start=1e-6, linear_end=1e-2)}
T_STEPS = 2000
# algorithmic work per view-image forward
GFLOP_PER_VIEW = 20.994
GFLOP_CONV_PER_VIEW = 2 * (9.595 + 0.723)        # conv3x3 + conv1x1 MACs -> FLOPs (attention core excluded)

def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel from the round's `ncu --set full` capture (scripts/ncu_traffic.py writes
    profiles/conv_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum averaged over the captured launches); None when
    no capture of this round is committed (ncu cannot run inside the timed bench)."""
    p = os.path.join(ROOT, "profiles", "conv_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None

# algorithmic HBM bytes of the conv class per view-image: every conv reads its sources once and writes its output once (bf16);
# weights are L2-resident.  From the layer table (Appendix A/B): sum over the 84 launches of (Cin_total + Cout) * H*W * 2 B.
CONV_ALGO_BYTES_PER_VIEW = 52.06e6
COMPOSE_BYTES_PER_SAMPLE = lambda n: n * 4096 * 32 + 2 * 3 * 4096 * 4


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def synthetic(B, N, seed=1234):
    """NMR-shaped synthetic inputs: CPU generator, then H2D."""
    g = torch.Generator().manual_seed(seed)
    y_cond = torch.rand(B, N, 3, 64, 64, generator=g)
    y_T = torch.randn(B, 3, 64, 64, generator=g)
    angle = (2 * torch.pi / 24) * torch.randint(0, 24, (B, 1), generator=g).float()
    view_count = torch.full((B,), N, dtype=torch.long)
    return y_cond, y_T, angle, view_count


# --------------------------------------------------------------------------------------------------
# CPU implementation (oracle port of the reference) — used ONLY as the reported baseline / reference arm
# --------------------------------------------------------------------------------------------------
def cpu_psample_rate(B, N, steps, warmup, threads):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import vf_oracle as O
    torch.set_num_threads(threads)
    sd = O.init_state_dict(O.SMALL_V100, 0, prefix="denoise_fn.")
    sched = O.make_schedule(**O.BETA_TRAIN)
    y_cond, y, angle, vc = synthetic(B, N)
    times = []
    with torch.no_grad():
        for j in range(warmup + steps):
            t = torch.full((B,), T_STEPS - 1 - j, dtype=torch.long)
            z = torch.randn(B, 3, 64, 64)
            t0 = time.perf_counter()
            y, *_ = O.p_sample(sd, O.SMALL_V100, sched, y, y_cond, vc, angle, t, z)
            if j >= warmup:
                times.append(time.perf_counter() - t0)
    sec = statistics.median(times)
    return B / (T_STEPS * sec), sec


def cpu_train_rate(B, N, steps, warmup, threads):
    """Reference-equivalent CPU training step (oracle forward + torch autograd backward + Adam), samples/s."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import vf_oracle as O
    torch.set_num_threads(threads)
    sd = O.init_state_dict(O.SMALL_V100, 0, prefix="denoise_fn.")
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    opt = torch.optim.Adam(list(params.values()), lr=1e-4)
    sched = O.make_schedule(**O.BETA_TRAIN)
    y_cond, noise, angle, vc = synthetic(B, N)
    g = torch.Generator().manual_seed(5)
    y0 = torch.rand(B, 3, 64, 64, generator=g)
    times = []
    for j in range(warmup + steps):
        t = torch.randint(1, T_STEPS, (B,), generator=g)
        u = torch.rand(B, 1, generator=g)
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss, _ = O.train_loss(params, O.SMALL_V100, sched, y0, y_cond, vc, angle, t, u, noise)
        loss.backward()
        opt.step()
        if j >= warmup:
            times.append(time.perf_counter() - t0)
    sec = statistics.median(times)
    return B / sec, sec



# --------------------------------------------------------------------------------------------------
# same-box LIBRARY baseline: the reference graph through torch's own CUDA kernels (cuDNN / cuBLAS / ATen)
# --------------------------------------------------------------------------------------------------
def library_baseline(B, N, dev, steps=5, warmup=2):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import vf_oracle as O
    out = {}
    sd0 = O.init_state_dict(O.SMALL_V100, 0, prefix="denoise_fn.")
    sched = {k: v.to(dev) for k, v in O.make_schedule(**O.BETA_TRAIN).items()}
    y_cond, y, angle, vc = synthetic(B, N)
    y_cond, y, angle = y_cond.to(dev), y.to(dev), angle.to(dev)
    z = torch.randn(B, 3, 64, 64, device=dev)
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.benchmark = True
    try:
        for mode in ("fp32_no_tf32", "bf16_autocast_channels_last"):
            bf = mode.startswith("bf16")
            torch.backends.cudnn.allow_tf32 = False
            torch.backends.cuda.matmul.allow_tf32 = False
            sd = {k: (v.to(dev).contiguous(memory_format=torch.channels_last) if (bf and v.dim() == 4) else v.to(dev)) for k, v in sd0.items()}
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=bf):
                yy = y
                for j in range(warmup + steps):
                    if j == warmup:
                        ev[0].record()
                    t = torch.full((B,), T_STEPS - 1 - j, dtype=torch.long, device=dev)
                    yy = O.p_sample(sd, O.SMALL_V100, sched, yy.float(), y_cond, vc, angle, t, z)[0]
                ev[1].record()
            torch.cuda.synchronize()
            ms = ev[0].elapsed_time(ev[1]) / steps
            out[mode] = {"ms_per_step": ms, "views_per_sec": B / (T_STEPS * ms * 1e-3)}
            del sd
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old
    torch.cuda.empty_cache()
    out["what"] = (f"reference graph (oracle/vf_oracle.py functional restatement) on this GPU through torch {torch.__version__} library kernels, "
                   f"B={B} N={N}, {steps} timed p_sample steps after {warmup} warm-up, cudnn.benchmark on; includes torch's host overhead")
    return out


def compose_roofline(model, dev, B=1024, N=6, iters=10):
    """The fused composition + DDPM kernel alone at a batch that fills the machine: algorithmic bytes / event time."""
    import ctypes as C
    from view_fusion_b200 import _lib
    lib = _lib.require_device()
    out8 = torch.randn(B * N * 4096, 8, device=dev)
    y_t, y_prev = torch.randn(B, 3, 64, 64, device=dev), torch.empty(B, 3, 64, 64, device=dev)
    off = (torch.arange(B + 1, dtype=torch.int32) * N).to(dev)
    t32 = torch.full((B,), 1000, dtype=torch.int32, device=dev)
    a = _lib.ComposeArgs()
    a.unet_out, a.view_offset, a.t, a.y_t, a.y_prev = out8.data_ptr(), off.data_ptr(), t32.data_ptr(), y_t.data_ptr(), y_prev.data_ptr()
    a.seed, a.offset, a.add_noise, a.clip_denoised, a.weighting, a.B, a.H, a.W, a.max_v = 1234, 1, 1, 1, 1, B, 64, 64, N
    sched = model._schedule_struct()
    st = _lib.stream_handle()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for j in range(iters + 3):
        if j == 3:
            ev[0].record()
        _lib.check(lib.vf_compose_ddpm_step(C.byref(a), C.byref(sched), st), "compose")
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / iters
    nbytes = B * COMPOSE_BYTES_PER_SAMPLE(N)
    pk = peaks()
    gbs = nbytes / (ms * 1e-3) / 1e9
    return {"kernel": "compose_ddpm_kernel", "B": B, "N": N, "bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s",
            "frac": gbs / pk["hbm"], "bytes_per_launch": nbytes, "avg_launch_ms": ms,
            "note": f"input ({B * N * 4096 * 32 / 1e6:.0f} MB) exceeds the 126 MB L2; fp32 [.,8] UNet output rows as produced by the final conv"}


def extra_configs(model, dev, steps, rank):
    """BASELINE.json configs 4 and 5: extrapolation N = 12 / 24 at B = 12 (experiment.py:472-514, vis batch size :212) and the
    autoregressive orbit at B = 1, count = 1..24 (experiment.py:516-578).  Per-step device time over `steps` consecutive
    reverse steps through generate(); the orbit time is the sum over the 24 growing conditioning sets of T * step(count)."""
    out = {}
    ev = lambda: torch.cuda.Event(enable_timing=True)
    host0 = None
    for n in (12, 24):
        y_cond, y_T, angle, vc = synthetic(12, n, seed=77 + n)
        y_cond, y_T, angle = y_cond.to(dev), y_T.to(dev), angle.to(dev)
        ts = [T_STEPS - 1 - j for j in range(steps)]
        model.generate(y_cond, vc, angle, y_t=y_T, steps=ts[:3])
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record()
        model.generate(y_cond, vc, angle, y_t=y_T, steps=ts)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out[f"extrapolation_N{n}_B12"] = {"ms_per_step": ms, "views_per_sec": 12 / (T_STEPS * ms * 1e-3), "view_images_per_step": 12 * n}
    # autoregressive: B = 1, the conditioning set grows from 1 to 24 views
    y_cond, y_T, angle, _ = synthetic(1, 24, seed=5)
    y_cond, y_T, angle = y_cond.to(dev), y_T.to(dev), angle.to(dev)
    per_count, host_ms = {}, {}
    total_s = 0.0
    ts = [T_STEPS - 1 - j for j in range(steps)]
    for count in range(1, 25):
        vc = torch.full((1,), count, dtype=torch.long)
        model.generate(y_cond, vc, angle, y_t=y_T, steps=ts[:3])
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        h0 = time.perf_counter()
        e0.record()
        model.generate(y_cond, vc, angle, y_t=y_T, steps=ts)
        e1.record()
        h1 = time.perf_counter()                  # host time to ENQUEUE the steps (no sync yet)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        per_count[count] = round(ms, 4)
        host_ms[count] = round((h1 - h0) * 1e3 / steps, 4)
        total_s += T_STEPS * ms * 1e-3
    out["autoregressive_B1"] = {"seconds_per_24_view_orbit": total_s, "ms_per_step_by_view_count": per_count,
                                "host_enqueue_ms_per_step_by_view_count": host_ms,
                                "how": f"sum over count = 1..24 of T * (device ms per step at that count, {steps} consecutive steps each); "
                                       "host_enqueue = wall time generate() needs to issue a step (must stay below the device time)"}
    return out


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    B = args.ref_batch
    vps, sec = cpu_psample_rate(B, args.views, max(1, args.steps), max(1, min(args.warmup, 2)), threads)
    line = {
        "impl": "reference", "metric": "sampled_views_per_sec", "value": vps, "unit": "views/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"small-v100 UNet 64x64 N={args.views} T={T_STEPS}; step = one p_sample reverse step on a "
                               f"B={B} sample of the B={args.batch} batch; views/s = B/(T*step)", "B": B, "N": args.views, "T": T_STEPS},
        "cpu_baseline": {"value": vps, "unit": "views/s", "cores": threads, "kind": "port",
                         "sample": f"oracle/vf_oracle.py p_sample, B={B} N={args.views}, median of {max(1, args.steps)} steps, full T extrapolated"},
        "e2e": {"value": vps, "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=28, help="samples per GPU (small-v100.yaml batch_size)")
    ap.add_argument("--views", type=int, default=6)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--ref-batch", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--train-batch", type=int, default=28, help="training samples per GPU (weak scaling)")
    ap.add_argument("--train-steps", type=int, default=20)
    ap.add_argument("--no-train", action="store_true", help="skip the training-throughput leg")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling training leg (global batch 160)")
    ap.add_argument("--global-batch", type=int, default=160, help="strong scaling: global batch (medium-a100-4.yaml:35)")
    ap.add_argument("--micro-batch", type=int, default=40, help="strong scaling: samples per forward/backward on one GPU")
    ap.add_argument("--no-full-generate", action="store_true", help="skip the real T=2000 generate() run (~13 s)")
    ap.add_argument("--no-library-baseline", action="store_true", help="skip the torch library-kernel comparison on this GPU")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the autoregressive (C4) / extrapolation (C5) legs")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch.distributed as dist
    from view_fusion_b200 import UNet, ViewFusion, _lib

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    torch.manual_seed(0)                                   # random-init weights: manual_seed(0) -> UNet -> ViewFusion
    with contextlib.redirect_stdout(io.StringIO()):
        model = ViewFusion(UNet(**SMALL, precision=args.precision), BETA)
    model = model.to(dev)
    model.set_new_noise_schedule(device=dev, phase="train")
    B, N = args.batch, args.views
    y_cond_h, y_T_h, angle_h, vc = synthetic(B, N, seed=1234 + rank)
    y_cond, y_t, angle = y_cond_h.to(dev), y_T_h.to(dev), angle_h.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(first_t, k, y_start):
        """k consecutive reverse steps t = first_t, first_t - 1, ... through the public generate() (wraps below 0)."""
        ts = [(first_t - j) % T_STEPS for j in range(k)]
        return model.generate(y_cond, vc, angle, y_t=y_start, steps=ts)[0]

    with torch.no_grad():
        y_w = run_steps(T_STEPS - 1, args.warmup, y_t)
        launches_per_step = model.denoise_fn.last_launches() + model.step_overhead_launches()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local_rank) as clk:
            e0.record()
            y_k = run_steps(T_STEPS - 1 - args.warmup, args.steps, y_w)
            e1.record()
            barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tm = torch.tensor([ms], device=dev)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ms = float(tm)
        ms_step = ms / args.steps
        value = world * B / (T_STEPS * ms_step * 1e-3)
        assert bool(torch.isfinite(y_k).all())

        # ---- one REAL full-T generate() through the module call the reference's experiment makes ------------------
        gen_full = None
        if not args.no_full_generate:
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            w0 = time.perf_counter()
            g0.record()
            y_fin, ret_arr, logit_arr, weight_arr, last = model(y_cond=y_cond, view_count=vc, angle=angle, generate=True)
            g1.record()
            barrier()
            wall = time.perf_counter() - w0
            gsec = g0.elapsed_time(g1) * 1e-3
            if world > 1:
                tm = torch.tensor([gsec, wall], device=dev)
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
                gsec, wall = float(tm[0]), float(tm[1])
            assert tuple(ret_arr.shape) == (B, 9, 3, 64, 64) and bool(torch.isfinite(y_fin).all())
            gen_full = {"value": world * B / gsec, "unit": "views/s", "seconds": gsec, "wall_seconds": wall, "T": T_STEPS,
                        "ms_per_step": gsec * 1e3 / T_STEPS, "ratio_to_k_step_value": (world * B / gsec) / value,
                        "api": "ViewFusion.forward(y_cond, view_count, angle, generate=True): T reverse steps + 8 weight/logit snapshots"}
            del y_fin, ret_arr, logit_arr, weight_arr, last

        # ---- e2e through the public API with host buffers -------------------------------------------------
        # Every step: y_cond / y_t / angle of THAT step come from pinned host memory (H2D), the step runs through the public
        # ViewFusion.p_sample, and y_{t-1} is read back to pinned host memory (D2H) and touched by the host.  The three phases
        # are software-pipelined the way a serving loop would: the inputs of step k+1 are copied on a second stream while step k
        # computes, and the host consumes the result of step k-1 while step k runs; nothing is skipped or cached.
        pin = lambda x: x.contiguous().pin_memory()
        yc_p, yt_p, an_p, vc_p = pin(y_cond_h), pin(y_T_h), pin(angle_h), pin(vc)
        t_host = torch.full((B,), T_STEPS - 1, dtype=torch.long).pin_memory()
        copy_st = torch.cuda.Stream(device=dev)
        main_st = torch.cuda.current_stream(dev)
        in_buf = [[torch.empty_like(y_cond), torch.empty_like(y_t), torch.empty_like(angle)] for _ in range(2)]
        out_buf = [torch.empty_like(yt_p).pin_memory() for _ in range(2)]
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_free = [torch.cuda.Event() for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(2)]
        host_sum = [0.0]

        def stage_inputs(k):
            b = k & 1
            with torch.cuda.stream(copy_st):
                copy_st.wait_event(ev_free[b])                     # the step that last used this buffer has consumed it
                in_buf[b][0].copy_(yc_p, non_blocking=True); in_buf[b][1].copy_(yt_p, non_blocking=True); in_buf[b][2].copy_(an_p, non_blocking=True)
                ev_in[b].record(copy_st)

        def e2e_loop(n):
            for b in range(2):
                ev_free[b].record(main_st)
            stage_inputs(0)
            for k in range(n):
                b = k & 1
                if k + 1 < n:
                    stage_inputs(k + 1)
                main_st.wait_event(ev_in[b])
                y_prev, _, _ = model.p_sample(in_buf[b][1], in_buf[b][0], vc_p, in_buf[b][2], t_host, want_weights=False)   # t, view_count: host tensors
                ev_free[b].record(main_st)
                out_buf[b].copy_(y_prev, non_blocking=True)
                ev_out[b].record(main_st)
                if k >= 1:                                          # the host consumes the previous step's result while this one runs
                    ev_out[b ^ 1].synchronize()
                    host_sum[0] += float(out_buf[b ^ 1][0, 0, 0, 0])
            ev_out[(n - 1) & 1].synchronize()
            host_sum[0] += float(out_buf[(n - 1) & 1][0, 0, 0, 0])

        e2e_loop(3)
        barrier()
        t0 = time.perf_counter()
        k_e2e = max(3, min(args.steps, 20))
        e2e_loop(k_e2e)
        barrier()
        e2e_s = (time.perf_counter() - t0) / k_e2e
        if world > 1:
            tm = torch.tensor([e2e_s], device=dev)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            e2e_s = float(tm)
        e2e_val = world * B / (T_STEPS * e2e_s)
        h2d = yc_p.numel() * 4 + yt_p.numel() * 4 + an_p.numel() * 4 + t_host.numel() * 8
        d2h = out_buf[0].numel() * 4
        assert host_sum[0] == host_sum[0]                          # finite: the host really read every result

        # ---- roofline of the dominant kernel class (one profiled step; rank 0) ------------------------------
        roof, classes = None, None
        if rank == 0:
            model.denoise_fn.set_profiling(True)
            acc = {}
            for j in range(3):
                run_steps(T_STEPS - 1 - j, 1, y_t)
                pr = model.denoise_fn.profile()
                if j > 0:
                    for k, (m_, c_) in pr.items():
                        a_ = acc.setdefault(k, [0.0, 0])
                        a_[0] += m_ / 2; a_[1] = c_
            model.denoise_fn.set_profiling(False)
            pk = peaks()
            tr_ = ncu_traffic()
            conv_ms, conv_n = acc["conv"]
            images = B * N
            flops = GFLOP_CONV_PER_VIEW * 1e9 * images
            ach = flops / (conv_ms * 1e-3) / 1e12
            roof = {"bound": "tensor", "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"],
                    "traffic": (tr_.get("bytes_per_launch") if (tr_ and (B, N) == (28, 6)) else None),
                    "traffic_source": (tr_.get("source") if tr_ else "no ncu capture committed for this round"),
                    "algorithmic_bytes_per_launch": CONV_ALGO_BYTES_PER_VIEW * images / conv_n,
                    "kernel": "conv_tc_kernel", "launches_per_step": conv_n,
                    "flops_per_launch": flops / conv_n, "avg_launch_ms": conv_ms / conv_n,
                    "peak_source": f"{pk['src']} sustained bf16 (kernel timed inside a long step)"}
            gn_bytes = 9_220_096 * images
            classes = {k: {"ms_per_step": round(v[0], 4), "launches": v[1]} for k, v in acc.items()}
            if acc["gn_apply"][0] > 0:
                classes["gn_apply"]["GBps_algorithmic_4B_per_elem"] = round(gn_bytes * 4 / (acc["gn_apply"][0] * 1e-3) / 1e9, 1)
            if acc["gn_stats"][0] > 0:
                classes["gn_stats"]["GBps_algorithmic_2B_per_elem"] = round(gn_bytes * 2 / (acc["gn_stats"][0] * 1e-3) / 1e9, 1)
            classes["hbm_peak_GBps"] = pk["hbm"]
            classes["step_flop_utilisation"] = round(GFLOP_PER_VIEW * 1e9 * images / (ms_step * 1e-3) / 1e12 / pk["tf_sustained"], 4)

    lib_base = roof_compose = extra = None
    if rank == 0 and world == 1:
        with torch.no_grad():
            roof_compose = compose_roofline(model, dev)
            if not args.no_extra_configs:
                extra = extra_configs(model, dev, max(10, min(args.steps, 40)), rank)
        model.release_buffers()
        torch.cuda.empty_cache()
        if not args.no_library_baseline:
            lib_base = library_baseline(B, N, dev)

    # ---- training leg: zero_grad -> forward -> backward (+ gradient all-reduce) -> Adam ------------------------
    train = None
    if not args.no_train:
        from view_fusion_b200.distributed import data_parallel
        torch.cuda.empty_cache()
        Bt = args.train_batch
        yc_h, eps_h, an_h, vct = synthetic(Bt, N, seed=4321 + rank)
        y0_h = torch.rand(Bt, 3, 64, 64, generator=torch.Generator().manual_seed(99 + rank))
        pin = lambda x: x.contiguous().pin_memory()
        yc_p, y0_p, an_p = pin(yc_h), pin(y0_h), pin(an_h)
        yc_d, y0_d, an_d = yc_h.to(dev), y0_h.to(dev), an_h.to(dev)
        if world > 1:
            data_parallel(model)
        from view_fusion_b200.optim import FusedAdam
        opt = FusedAdam(model.parameters(), lr=1e-4)         # torch.optim.Adam semantics, one launch

        def train_step(host: bool):
            if host:
                yc, y0, an = yc_p.to(dev, non_blocking=True), y0_p.to(dev, non_blocking=True), an_p.to(dev, non_blocking=True)
            else:
                yc, y0, an = yc_d, y0_d, an_d
            opt.zero_grad(set_to_none=True)
            loss = model(y_cond=yc, view_count=vct, angle=an, y_0=y0)
            loss.backward()
            opt.step()
            return float(loss.detach()) if host else loss.detach()

        for _ in range(3):
            train_step(False)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.train_steps):
            last_loss = train_step(False)
        e1.record()
        barrier()
        tr_ms = e0.elapsed_time(e1) / args.train_steps
        train_step(True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.train_steps):
            last = train_step(True)
        barrier()
        tr_e2e = (time.perf_counter() - t0) / args.train_steps
        if world > 1:
            tm = torch.tensor([tr_ms, tr_e2e], device=dev)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            tr_ms, tr_e2e = float(tm[0]), float(tm[1])
        pk = peaks()
        tr_flops = 3 * GFLOP_PER_VIEW * 1e9 * Bt * N
        train = {
            "metric": "train_samples_per_sec", "value": world * Bt / (tr_ms * 1e-3), "unit": "samples/s", "ms_per_step": tr_ms,
            "steps": args.train_steps, "scaling": "weak", "B_per_gpu": Bt, "N": N,
            "step": "zero_grad -> ViewFusion.forward (loss) -> backward (hand-written CUDA) -> "
                    + ("[NCCL all-reduce (AVG) of each of 4 backward phases' gradient slice, issued as the phase is enqueued and "
                       "overlapped with the remaining phases] -> " if world > 1 else "") + "fused Adam (vf_adam_step)",
            "e2e": {"value": world * Bt / tr_e2e, "unit": "samples/s", "ms_per_step": tr_e2e * 1e3,
                    "h2d_bytes_per_step": (yc_p.numel() + y0_p.numel() + an_p.numel()) * 4, "d2h_bytes_per_step": 4},
            "flop_utilisation": round(tr_flops / (tr_ms * 1e-3) / 1e12 / pk["tf_sustained"], 4),
            "loss": float(last),
        }
        # ---- strong scaling (BASELINE config C3, medium-a100-4.yaml:35): GLOBAL batch 160 split over the ranks; a rank's share
        # runs as micro-batches of <= 40 samples whose gradients accumulate natively in the flat buffer, and only the last
        # micro-batch's backward exchanges them (no_sync on the others)
        if not args.no_strong:
            import contextlib as _cl
            from view_fusion_b200.distributed import no_sync
            G = args.global_batch
            per = G // world
            micro = min(args.micro_batch, per)
            n_micro = per // micro
            per = micro * n_micro
            ycs, _, ans, vcs = synthetic(per, N, seed=777 + rank)
            y0s = torch.rand(per, 3, 64, 64, generator=torch.Generator().manual_seed(55 + rank))
            ycs, y0s, ans = ycs.to(dev), y0s.to(dev), ans.to(dev)
            model.denoise_fn.grad_accumulation(True)

            def strong_step():
                opt.zero_grad(set_to_none=True)
                tot = None
                for m in range(n_micro):
                    sl = slice(m * micro, (m + 1) * micro)
                    ctx = no_sync(model) if (world > 1 and m < n_micro - 1) else _cl.nullcontext()
                    with ctx:
                        loss = model(y_cond=ycs[sl], view_count=vcs[sl], angle=ans[sl], y_0=y0s[sl]) * (1.0 / n_micro)
                        loss.backward()
                    tot = loss.detach() if tot is None else tot + loss.detach()
                opt.step()
                return tot

            for _ in range(2):
                strong_step()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ks = max(3, args.train_steps // 2)
            e0.record()
            for _ in range(ks):
                ls = strong_step()
            e1.record()
            barrier()
            st_ms = e0.elapsed_time(e1) / ks
            if world > 1:
                tm = torch.tensor([st_ms], device=dev)
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
                st_ms = float(tm)
            model.denoise_fn.grad_accumulation(False)
            train["strong"] = {"metric": "train_samples_per_sec", "scaling": "strong", "global_batch": per * world, "B_per_gpu": per,
                               "micro_batch": micro, "micro_batches_per_step": n_micro, "ms_per_step": st_ms,
                               "value": per * world / (st_ms * 1e-3), "steps": ks, "loss": float(ls),
                               "flop_utilisation": round(3 * GFLOP_PER_VIEW * 1e9 * per * N / (st_ms * 1e-3) / 1e12 / pk["tf_sustained"], 4)}
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            sps, sec = cpu_train_rate(2, N, steps=2, warmup=1, threads=threads)
            train["cpu_baseline"] = {"value": sps, "unit": "samples/s", "cores": threads, "kind": "port",
                                     "sample": f"oracle train step (fwd + torch autograd bwd + Adam), B=2 N={N}: 1 warm-up + 2 timed, "
                                               f"median {sec:.2f} s/step"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        vps, sec = cpu_psample_rate(2, N, steps=3, warmup=1, threads=threads)
        cpu = {"value": vps, "unit": "views/s", "cores": threads, "kind": "port",
               "sample": f"oracle p_sample (CPU fp32 restatement of the reference), B=2 N={N}: 1 warm-up + 3 timed steps, "
                         f"median {sec:.3f} s/step, full T={T_STEPS} extrapolated"}

    if rank == 0:
        line = {
            "metric": "sampled_views_per_sec", "value": value, "unit": "views/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": f"small-v100: UNet 33.9M params, 64x64, B={B}/GPU, N={N} views ({B * N} view-images/step), "
                                   f"T={T_STEPS}; step = one p_sample reverse step; views/s = n_gpus*B/(T*step)",
                       "B_per_gpu": B, "N": N, "T": T_STEPS, "parallelism": f"sample-batch sharded x{world}, no collective",
                       "l2": "per-step activation working set (>10 GB) far exceeds the 126 MB L2; no flush needed"},
            "e2e": {"value": e2e_val, "unit": "views/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "ViewFusion.p_sample(host pinned tensors) -> host y_{t-1}; H2D of step k+1 and the host read of step k-1 "
                           "overlap step k (two streams, double-buffered)", "ms_per_step": e2e_s * 1e3},
            "gpu_launches": launches_per_step * args.steps,
            "gpu_launches_per_step": launches_per_step,
            "clocks": clk.summary(),
            "roofline": roof,
            "roofline_compose": roof_compose,
            "generate_full": gen_full,
            "library_baseline": lib_base,
            "extra_configs": extra,
            "kernel_classes": classes,
            "cpu_baseline": cpu,
            "train": train,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

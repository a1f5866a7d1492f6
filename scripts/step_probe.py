"""Sampling step: kernel-time sum vs wall time (gaps), per-kernel table via torch profiler."""
import contextlib, io, sys, time
import torch
sys.path.insert(0, ".")
from view_fusion_b200 import UNet, ViewFusion
from view_fusion_b200.view_fusion import _Plan
from bench import SMALL, BETA, synthetic

B, N = 28, 6
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    model = ViewFusion(UNet(**SMALL, precision="bf16"), BETA).cuda()
model.set_new_noise_schedule(device="cuda", phase="train")
y_cond, y_T, angle, vc = synthetic(B, N)
y_cond, y_t, angle = y_cond.cuda(), y_T.cuda(), angle.cuda()
plan = _Plan(model, y_cond, vc)
bufs = [y_t, torch.empty_like(y_t)]
t = torch.empty(B, dtype=torch.long, device="cuda")

def step(j):
    t.fill_(1999 - j)
    model._step(plan, bufs[j & 1], y_cond, angle, t, bufs[(j + 1) & 1], add_noise=True)

with torch.no_grad():
    for j in range(5):
        step(j)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for j in range(20):
        step(j)
    e1.record()
    torch.cuda.synchronize()
    print(f"step {e0.elapsed_time(e1) / 20:.3f} ms")
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for j in range(4):
            step(j)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    tot = sum(e.device_time for e in evs) / 4
    print(f"kernel-time sum per step {tot / 1e3:.3f} ms over {len(evs) // 4} launches")
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=10, max_name_column_width=50))

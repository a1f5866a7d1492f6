"""Per-launch table of one profiled sampling step (events around every launch, PDL suspended): time, TFLOP/s or GB/s."""
import contextlib, io, sys
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
    model.denoise_fn.set_profiling(True)
    acc = None
    REP = 5
    for j in range(REP):
        step(j)
        torch.cuda.synchronize()
        rows = model.denoise_fn.profile_launches()
        if acc is None:
            acc = [list(r) for r in rows]
            for r in acc:
                r[1] = [r[1]]
        else:
            for a, r in zip(acc, rows):
                a[1].append(r[1])
    model.denoise_fn.set_profiling(False)

tot = {}
print(f"{'#':>3s} {'class':9s} {'img':>4s} {'H':>3s} {'K/C':>5s} {'Cout':>4s} k s {'us':>8s} {'TFLOP/s|GB/s':>12s}")
for i, (kind, ms, img, H, cin, cout, ks, st) in enumerate(acc):
    us = sorted(ms)[len(ms) // 2] * 1e3
    rate = ""
    if kind == "conv" and H:
        Ho = H // max(st, 1) if st else H
        fl = 2.0 * img * Ho * Ho * cin * cout
        rate = f"{fl / us / 1e6:9.1f} TF"
    elif kind == "gn_apply" and H:
        by = 4.0 * img * H * H * cin
        rate = f"{by / us / 1e3:9.1f} GB"
    print(f"{i:3d} {kind:9s} {img:4d} {H:3d} {cin:5d} {cout:4d} {ks} {st} {us:8.1f} {rate:>12s}")
    tot.setdefault(kind, [0.0, 0])
    tot[kind][0] += us; tot[kind][1] += 1
print({k: (round(v[0] / 1e3, 3), v[1]) for k, v in tot.items()})

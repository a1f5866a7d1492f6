"""A few training steps at the benchmark size (B=28, N=6, bf16) for an ncu launch list:
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_train.csv python scripts/train_launches.py"""
import contextlib, io, sys
import torch
sys.path.insert(0, ".")
from view_fusion_b200 import UNet, ViewFusion
from view_fusion_b200.optim import FusedAdam
from bench import SMALL, BETA, synthetic

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B, N = 28, 6
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    model = ViewFusion(UNet(**SMALL, precision="bf16"), BETA).cuda()
model.set_new_noise_schedule(device="cuda", phase="train")
y_cond, _, angle, vc = synthetic(B, N, seed=4321)
y0 = torch.rand(B, 3, 64, 64, generator=torch.Generator().manual_seed(99))
y_cond, y0, angle = y_cond.cuda(), y0.cuda(), angle.cuda()
opt = FusedAdam(model.parameters(), lr=1e-4)
for _ in range(steps):
    opt.zero_grad(set_to_none=True)
    loss = model(y_cond=y_cond, view_count=vc, angle=angle, y_0=y0)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
print("loss", float(loss.detach()))

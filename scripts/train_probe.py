"""Times one training step (zero_grad -> forward -> backward -> Adam) at the benchmark shape; prints a per-kernel breakdown via torch profiler."""
import contextlib, io, sys, time
import torch
sys.path.insert(0, ".")
from view_fusion_b200 import UNet, ViewFusion
from bench import SMALL, BETA, synthetic

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N = 6
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    model = ViewFusion(UNet(**SMALL, precision="bf16"), BETA)
model = model.cuda()
model.set_new_noise_schedule(device="cuda", phase="train")
opt = torch.optim.Adam(model.parameters(), lr=1e-4)
y_cond, y0, angle, vc = synthetic(B, N)
y_cond, y0, angle = y_cond.cuda(), torch.rand(B, 3, 64, 64, device="cuda"), angle.cuda()

def step():
    opt.zero_grad(set_to_none=True)
    loss = model(y_cond=y_cond, view_count=vc, angle=angle, y_0=y0)
    loss.backward()
    opt.step()
    return loss

for _ in range(2):
    l = step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    l = step()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
print(f"B={B}: {dt*1e3:.1f} ms/step, {B/dt:.1f} samples/s, loss {float(l):.4f}")
if len(sys.argv) > 2 and sys.argv[2] == "noprof":
    sys.exit(0)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=60))

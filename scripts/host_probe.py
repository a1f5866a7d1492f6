"""Host-side enqueue cost of the training step pieces (is the step launch-bound?)."""
import contextlib, io, sys, time
import torch
sys.path.insert(0, ".")
from view_fusion_b200 import UNet, ViewFusion
from bench import SMALL, BETA, synthetic

B, N = 28, 6
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    model = ViewFusion(UNet(**SMALL, precision="bf16"), BETA).cuda()
model.set_new_noise_schedule(device="cuda", phase="train")
opt = torch.optim.Adam(model.parameters(), lr=1e-4)
y_cond, y0, angle, vc = synthetic(B, N)
y_cond, y0, angle = y_cond.cuda(), torch.rand(B, 3, 64, 64, device="cuda"), angle.cuda()
def step(timing=None):
    t0 = time.perf_counter()
    opt.zero_grad(set_to_none=True)
    t1 = time.perf_counter()
    loss = model(y_cond=y_cond, view_count=vc, angle=angle, y_0=y0)
    t2 = time.perf_counter()
    loss.backward()
    t3 = time.perf_counter()
    opt.step()
    t4 = time.perf_counter()
    if timing is not None:
        timing.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3))
for _ in range(3):
    step()
torch.cuda.synchronize()
tm = []
for _ in range(5):
    torch.cuda.synchronize()          # empty queue: enqueue times are pure host cost
    step(tm)
torch.cuda.synchronize()
import statistics
for i, name in enumerate(("zero_grad", "forward enqueue", "backward enqueue", "Adam enqueue")):
    print(f"{name:18s} {statistics.median(t[i] for t in tm) * 1e3:7.2f} ms")
print(f"total host         {statistics.median(sum(t) for t in tm) * 1e3:7.2f} ms")

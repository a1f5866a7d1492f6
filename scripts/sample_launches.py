"""A few EAGER sampling steps at the benchmark size (B=28, N=6, bf16) for ncu launch lists / captures."""
import contextlib, io, sys
import torch
sys.path.insert(0, ".")
from view_fusion_b200 import UNet, ViewFusion
from bench import SMALL, BETA, synthetic

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B, N = 28, 6
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    model = ViewFusion(UNet(**SMALL, precision="bf16"), BETA).cuda()
model.set_new_noise_schedule(device="cuda", phase="train")
model.use_cuda_graph = False
y_cond, y_T, angle, vc = synthetic(B, N)
y = model.generate(y_cond.cuda(), vc, angle.cuda(), y_t=y_T.cuda(), steps=[1999 - j for j in range(steps)])[0]
torch.cuda.synchronize()
print("ok", bool(torch.isfinite(y).all()), model.denoise_fn.last_launches() + model.step_overhead_launches(), "launches per step")

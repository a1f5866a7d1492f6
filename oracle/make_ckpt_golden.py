"""Writes tests/golden/ref_checkpoint_micro.pt with the UNMODIFIED reference code (run in the build container only):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_ckpt_golden.py

The reference's own modules (model/unet.py, model/view_fusion.py), its own `utils.checkpoint.Checkpoint` (utils/checkpoint.py:31-47)
and `utils.schedulers.LrScheduler` build a tiny ViewFusion, take two real `torch.optim.Adam` steps at the scheduler's learning rate
on a CPU training loss, and save exactly the way experiment.py:121-128 / :242-254 do: Checkpoint(model=model_module,
optimizer=optimizer).save("model.pt", it=, t=, run_id=, <best metrics>).  Alongside, a few forward outputs of the SAVED weights are stored so the
loader test can check that the drop-in module computes the same function after loading (tests/test_interop_cpu.py, tests/test_gpu_configs.py).

TEST INFRASTRUCTURE: nothing in view_fusion_b200/ reads this script or /root/reference.
"""
import contextlib
import io
import os
import shutil
import sys
import tempfile

import numpy as np
import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REF)
from model.unet import UNet as RefUNet                      # noqa: E402
from model.view_fusion import ViewFusion as RefViewFusion   # noqa: E402
from utils.checkpoint import Checkpoint                     # noqa: E402
from utils.schedulers import LrScheduler                    # noqa: E402

MICRO = dict(in_channel=6, out_channel=6, inner_channel=32, norm_groups=32, channel_mults=(1,), attn_res=(16,), res_blocks=1, dropout=0,
             image_size=16)
BETA = {"train": dict(schedule="linear", n_timestep=2000, linear_start=1e-6, linear_end=1e-2)}


def main():
    torch.manual_seed(11)
    import inspect
    beta = BETA
    if "num_timesteps" in inspect.signature(sys.modules["model.view_fusion"].make_beta_schedule).parameters:
        beta = {"train": dict(schedule="linear", num_timesteps=2000, linear_start=1e-6, linear_end=1e-2)}
    with contextlib.redirect_stdout(io.StringIO()):
        model = RefViewFusion(RefUNet(**MICRO), beta)
    model.set_new_noise_schedule(device="cpu", phase="train")
    sched = LrScheduler(peak_lr=1e-3, peak_it=4, decay_rate=0.5, decay_it=10)
    opt = torch.optim.Adam(model.parameters(), lr=sched.get_cur_lr(0))
    g = torch.Generator().manual_seed(3)
    B, N, S = 2, 3, 16
    lrs = []
    for it in range(1, 3):
        for pg in opt.param_groups:                         # experiment.py:265-267
            pg["lr"] = sched.get_cur_lr(it)
        lrs.append(sched.get_cur_lr(it))
        y_cond = torch.rand(B, N, 3, S, S, generator=g)
        y_0 = torch.rand(B, 3, S, S, generator=g)
        angle = torch.rand(B, 1, generator=g)
        vc = torch.tensor([3, 2])
        opt.zero_grad()
        loss = model(y_0=y_0, y_cond=y_cond, view_count=vc, angle=angle)
        loss.backward()
        opt.step()
    tmp = tempfile.mkdtemp()
    ck = Checkpoint(os.path.join(tmp, "run"), device="cpu", rank=0, config={"model": MICRO}, model=model, optimizer=opt)
    with contextlib.redirect_stdout(io.StringIO()):
        ck.save("model.pt", it=2, t=1.5, run_id="ref-run", best_psnr=17.25)
    dst = os.path.join(ROOT, "tests", "golden", "ref_checkpoint_micro.pt")
    shutil.copyfile(os.path.join(tmp, "run", "model.pt"), dst)
    # reference outputs of the SAVED weights (UNet forward on fixed inputs) for the functional check after loading
    x = torch.randn(3, 6, S, S, generator=g)
    ang = torch.rand(3, 1, generator=g)
    lvl = torch.rand(3, 1, generator=g) * 0.98 + 0.01
    with torch.no_grad():
        out = model.denoise_fn(x, ang, lvl)
    np.savez(os.path.join(ROOT, "tests", "golden", "ref_checkpoint_micro_io.npz"), x=x.numpy(), angle=ang.numpy(), level=lvl.numpy(), out=out.numpy(),
             lrs=np.array(lrs), lr_probe_it=np.array([0, 1, 3, 4, 9, 14, 104]), lr_probe=np.array([sched.get_cur_lr(i) for i in (0, 1, 3, 4, 9, 14, 104)]),
             lr_args=np.array([1e-3, 4, 0.5, 10]))
    shutil.rmtree(tmp)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()

"""Helpers shared by the -m gpu tests."""
import contextlib
import io
import os

import numpy as np
import torch

import vf_oracle as O

BETA = {"train": dict(O.BETA_TRAIN)}
# bf16-capable toy config (every level a multiple of 64 channels; attention at L=64 and in mid at L=16 is avoided)
TOY64 = dict(in_channel=6, out_channel=6, inner_channel=64, norm_groups=32, channel_mults=(1, 2), attn_res=(8,),
             res_blocks=1, image_size=16)


def load(golden_dir, name):
    d = np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False)
    return {k: (torch.from_numpy(d[k]) if d[k].dtype.kind in "fi" and d[k].ndim > 0 else d[k]) for k in d.files}


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def build_model(cfg, seed, precision, weighting=True):
    from view_fusion_b200 import UNet, ViewFusion
    with contextlib.redirect_stdout(io.StringIO()):
        m = ViewFusion(UNet(**cfg, precision=precision), BETA, weighting_train=weighting, weighting_inference=weighting)
    m.set_new_noise_schedule(device="cpu", phase="train")
    sd = O.init_state_dict(cfg, seed, prefix="denoise_fn.")
    sd.update(O.make_schedule(**O.BETA_TRAIN))
    m.load_state_dict(sd, strict=True)
    return m.cuda(), {k: v for k, v in sd.items()}


def bf16r(x):
    return x.to(torch.bfloat16).float()


# ---- recorded margins: every parity number the GPU tests measure goes to one text file (copied to profiles/ per round) ----
_MARGIN_FILE = os.environ.get("VF_MARGINS_FILE") or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                                 "gpurun_out", "parity_margins.txt")


def margin(name, value, bar, higher_is_better=False):
    """Append `name value bar margin` to the margins file and return whether the bar is met (NaN never passes)."""
    value = float(value)
    ok = (value > bar) if higher_is_better else (value < bar)
    ratio = (value / bar) if higher_is_better else (bar / value if value > 0 else float("inf"))
    try:
        os.makedirs(os.path.dirname(_MARGIN_FILE), exist_ok=True)
        with open(_MARGIN_FILE, "a") as fh:
            fh.write(f"{name:<78s} measured {value:11.4e}  bar {'>' if higher_is_better else '<'} {bar:9.3e}  margin x{ratio:7.2f}  {'ok' if ok else 'FAIL'}\n")
    except OSError:
        pass
    return ok


def argmax_agrees(w, w_ref, gap):
    """View-weight argmax (over the view axis, per pixel and channel) identical wherever the reference's two largest weights
    differ by more than `gap`.  Random-init logits are near-ties: fp32 mode uses gap = 1e-5 (a few pixels are
    ties below fp32 noise, where two fp32 evaluations — or two runs of the atomically accumulated GroupNorm sums — disagree),
    bf16 mode 2e-2.  Returns (all decided pixels agree, fraction of ALL pixels that agree)."""
    w, w_ref = w.detach().float().cpu(), w_ref.detach().float().cpu()
    same = w.argmax(1) == w_ref.argmax(1)
    if w_ref.shape[1] < 2:
        return bool(same.all()), 1.0
    top2 = w_ref.topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > gap
    return bool(same[clear].all()), float(same.float().mean())

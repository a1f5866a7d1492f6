"""Reference-format checkpoint files (utils/checkpoint.py:31-72) load into the drop-in modules and back."""
import contextlib
import io
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import vf_oracle as O

from view_fusion_b200 import UNet, ViewFusion
from view_fusion_b200.interop import load_checkpoint, save_checkpoint
from view_fusion_b200.optim import FusedAdam

BETA = {"train": dict(O.BETA_TRAIN)}


def _model(seed):
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        m = ViewFusion(UNet(**O.TINY, precision="fp32"), BETA)
    m.set_new_noise_schedule(device="cpu", phase="train")
    return m


def _reference_style_file(path, ddp_prefix=False):
    """What experiment.py:242-254 writes: the reference model's state_dict (here: the oracle's deterministic weights under
    the reference's key names) + torch.optim.Adam's state_dict + bookkeeping."""
    sd = O.init_state_dict(O.TINY, 7, prefix="denoise_fn.")
    sd.update(O.make_schedule(**O.BETA_TRAIN))
    donor = _model(1)
    donor.load_state_dict(sd, strict=True)
    opt = torch.optim.Adam(donor.parameters(), lr=3e-5)
    for p in donor.parameters():
        p.grad = torch.full_like(p, 0.01)
    opt.step()                                            # gives every parameter exp_avg / exp_avg_sq / step
    model_sd = donor.state_dict()
    if ddp_prefix:
        model_sd = {"module." + k: v for k, v in model_sd.items()}
    torch.save({"model": model_sd, "optimizer": opt.state_dict(), "it": 41, "t": 12.5, "run_id": "abc", "psnr": 20.0}, path)
    return donor, opt


@pytest.mark.parametrize("ddp_prefix", [False, True])
def test_reference_checkpoint_loads_strict(tmp_path, ddp_prefix):
    path = str(tmp_path / "best_model_all.pt")
    donor, donor_opt = _reference_style_file(path, ddp_prefix)
    m = _model(2)
    opt = FusedAdam(m.parameters(), lr=1e-4)
    rest = load_checkpoint(path, m, opt, map_location="cpu")
    assert rest == {"it": 41, "t": 12.5, "run_id": "abc", "psnr": 20.0}
    for (k, a), (_, b) in zip(m.state_dict().items(), donor.state_dict().items()):
        assert torch.equal(a, b), k
    assert opt.param_groups[0]["lr"] == 3e-5                       # the scheduler's value travels with the file
    for p, q in zip(m.parameters(), donor.parameters()):
        for key in ("exp_avg", "exp_avg_sq"):
            assert torch.equal(opt.state[p][key], donor_opt.state[q][key])
        assert float(opt.state[p]["step"]) == 1.0


def test_save_checkpoint_is_readable_by_the_reference_pattern(tmp_path):
    m = _model(3)
    opt = FusedAdam(m.parameters(), lr=2e-4)
    path = str(tmp_path / "model.pt")
    save_checkpoint(path, m, opt, it=7, t=1.0, run_id=None)
    raw = torch.load(path, map_location="cpu", weights_only=False)
    assert set(raw) == {"model", "optimizer", "it", "t", "run_id"}
    assert len(raw["model"]) == 6 + len(O.init_state_dict(O.TINY, 0))       # schedule buffers + UNet entries, reference key names
    assert all(k.startswith("denoise_fn.") or k in O.make_schedule(**O.BETA_TRAIN) for k in raw["model"])
    # the reference's loader: module.load_state_dict(state_dict[k]) for k in ("model", "optimizer") — utils/checkpoint.py:63-66
    m2 = _model(4)
    m2.load_state_dict(raw["model"])
    torch.optim.Adam(m2.parameters(), lr=1.0).load_state_dict(raw["optimizer"])
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))


def test_missing_model_entry_is_an_error(tmp_path):
    path = str(tmp_path / "x.pt")
    torch.save({"optimizer": {}}, path)
    with pytest.raises(KeyError):
        load_checkpoint(path, _model(5))


# ---- a file written by the reference's OWN code (oracle/make_ckpt_golden.py: model/unet.py + model/view_fusion.py +
# utils/checkpoint.py:Checkpoint.save + torch.optim.Adam + utils/schedulers.py:LrScheduler, run unmodified) -------------------
MICRO = dict(in_channel=6, out_channel=6, inner_channel=32, norm_groups=32, channel_mults=(1,), attn_res=(16,), res_blocks=1, image_size=16)
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_file_written_by_the_reference_checkpoint_class_loads_strict():
    import numpy as np
    from view_fusion_b200.optim import LrScheduler
    torch.manual_seed(99)
    with contextlib.redirect_stdout(io.StringIO()):
        m = ViewFusion(UNet(**MICRO, precision="fp32"), BETA)
    m.set_new_noise_schedule(device="cpu", phase="train")
    opt = FusedAdam(m.parameters(), lr=1.0)
    raw = torch.load(os.path.join(GOLD, "ref_checkpoint_micro.pt"), map_location="cpu", weights_only=False)
    assert set(raw) == {"model", "optimizer", "it", "t", "run_id", "best_psnr"}
    assert list(raw["model"]) == list(m.state_dict()), "state_dict key order of the drop-in == the reference module's"
    rest = load_checkpoint(os.path.join(GOLD, "ref_checkpoint_micro.pt"), m, opt, map_location="cpu")     # strict=True
    assert rest == {"it": 2, "t": 1.5, "run_id": "ref-run", "best_psnr": 17.25}
    for k, v in m.state_dict().items():
        assert torch.equal(v, raw["model"][k]), k
    io_ = np.load(os.path.join(GOLD, "ref_checkpoint_micro_io.npz"))
    assert abs(opt.param_groups[0]["lr"] - float(io_["lrs"][-1])) < 1e-12      # the scheduler's last value travels with the file
    n_state = 0
    for i, p in enumerate(m.parameters()):
        st, ref = opt.state[p], raw["optimizer"]["state"][i]
        assert float(st["step"]) == 2.0 and torch.equal(st["exp_avg"], ref["exp_avg"]) and torch.equal(st["exp_avg_sq"], ref["exp_avg_sq"])
        n_state += 1
    assert n_state == len(raw["optimizer"]["state"])
    # LrScheduler parity (utils/schedulers.py:1-14) on the probe iterations the reference evaluated
    a = io_["lr_args"]
    s = LrScheduler(peak_lr=float(a[0]), peak_it=int(a[1]), decay_rate=float(a[2]), decay_it=int(a[3]))
    for it, want in zip(io_["lr_probe_it"], io_["lr_probe"]):
        assert s.get_cur_lr(int(it)) == float(want)
    assert s.apply(opt, 9) == opt.param_groups[0]["lr"] == float(io_["lr_probe"][4])

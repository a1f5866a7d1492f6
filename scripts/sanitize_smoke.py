"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):
one bf16 sampling step (eager + CUDA-graph replay), one bf16 training step (forward, hand-written backward, phased backward,
fused Adam) and one fp32 (CUDA-core) sampling + training step, on the 16x16 toy configuration.

    compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py
"""
import contextlib, io, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import vf_oracle as O
from view_fusion_b200 import UNet, ViewFusion
from view_fusion_b200.optim import FusedAdam

TOY64 = dict(in_channel=6, out_channel=6, inner_channel=64, norm_groups=32, channel_mults=(1, 2), attn_res=(8,), res_blocks=1, image_size=16)
BETA = {"train": dict(O.BETA_TRAIN)}
graphs = os.environ.get("VF_SANITIZE_GRAPHS", "1") == "1"
for prec, cfg in (("bf16", TOY64), ("fp32", O.TINY)):
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = ViewFusion(UNet(**cfg, precision=prec), BETA).cuda()
    m.set_new_noise_schedule(device="cuda", phase="train")
    m.use_cuda_graph = graphs
    S = cfg["image_size"]
    g = torch.Generator().manual_seed(1)
    y_cond = torch.rand(3, 4, 3, S, S, generator=g).cuda()
    angle = torch.rand(3, 1, generator=g).cuda()
    vc = torch.tensor([4, 1, 3])
    y, ret, la, wa, last = m.generate(y_cond, vc, angle, steps=[1999, 1998, 1997, 1996, 1750, 1, 0])
    torch.cuda.synchronize()
    assert torch.isfinite(y).all()
    opt = FusedAdam(m.parameters(), lr=1e-4)
    y0 = torch.rand(3, 3, S, S, generator=g).cuda()
    for phased in (False, True):
        m.denoise_fn._grad_sync = (None, 3) if phased else None
        m.denoise_fn._layout_cache = None
        opt.zero_grad(set_to_none=True)
        loss = m(y_cond=y_cond, view_count=vc, angle=angle, y_0=y0)
        loss.backward()
        opt.step()
        torch.cuda.synchronize()
        assert torch.isfinite(loss)
    print(prec, "ok", float(loss))

"""Stand-alone timing of the GroupNorm forward / backward operators on benchmark-sized layers (168 view-images, bf16):
CUDA events around `iters` back-to-back calls, L2 flushed by the working set itself (>= 2 x 90 MB on the large layers)."""
import os, sys
import ctypes as C
import torch
sys.path.insert(0, ".")
from view_fusion_b200 import _lib, ops

lib = _lib.require_device()
R = 168
shapes = [(64, 0, 64), (128, 64, 64), (128, 0, 32), (192, 128, 32), (192, 0, 16), (192, 192, 16), (320, 0, 8), (320, 320, 8)]
if os.environ.get("GN_SHAPES"):
    shapes = shapes[: int(os.environ["GN_SHAPES"])]
iters = int(os.environ.get("GN_ITERS", "20"))
tag = os.environ.get("GN_TAG", "")
tot_f = tot_b = 0.0
for C0, C1, S in shapes:
    P = (S + 1) * (S + 1)
    x0 = torch.randn(R * P, C0, device="cuda").bfloat16()
    x1 = torch.randn(R * P, C1, device="cuda").bfloat16() if C1 else None
    st = ops.gn_stats(x0, x1, R, S, S)
    Cc = C0 + C1
    gm, bt = torch.rand(Cc, device="cuda") + 0.5, torch.randn(Cc, device="cuda") * 0.1
    dst = torch.empty(R * P, Cc, device="cuda", dtype=torch.bfloat16)
    dy = torch.randn(R * P, Cc, device="cuda").bfloat16()
    dx0, dx1 = torch.empty_like(x0), (torch.empty_like(x1) if C1 else None)
    scratch = torch.zeros(R * Cc * 2, device="cuda")
    dg, db = torch.zeros(Cc, device="cuda"), torch.zeros(Cc, device="cuda")
    sp = lambda t: 0 if t is None else t.data_ptr()

    def fwd():
        _lib.check(lib.vf_gn_apply(x0.data_ptr(), C0, st.data_ptr(), Cc, sp(x1), C1, st.data_ptr() + 8 * C0 if C1 else 0, Cc, _lib.VF_BF16, R, S, S, 32,
                                   gm.data_ptr(), bt.data_ptr(), 1, dst.data_ptr(), None, _lib.stream_handle()), "fwd")

    def bwd():
        _lib.check(lib.vf_gn_backward(x0.data_ptr(), C0, st.data_ptr(), Cc, sp(x1), C1, st.data_ptr() + 8 * C0 if C1 else 0, Cc, _lib.VF_BF16, R, S, S, 32,
                                      gm.data_ptr(), bt.data_ptr(), 1, dy.data_ptr(), scratch.data_ptr(), dg.data_ptr(), db.data_ptr(), dx0.data_ptr(), 0,
                                      sp(dx1), 0, None, _lib.stream_handle()), "bwd")

    res = []
    for fn in (fwd, bwd):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / iters * 1e3)
    elems = R * S * S * Cc
    tot_f += res[0]; tot_b += res[1]
    print(f"{tag:14s} C={C0}+{C1} {S}x{S}: fwd {res[0]:7.1f} us ({elems * 4 / res[0] / 1e3:6.0f} GB/s of 4 B/elem)   bwd {res[1]:7.1f} us ({elems * 10 / res[1] / 1e3:6.0f} GB/s of 10 B/elem)")
print(f"{tag:14s} total fwd {tot_f:7.1f} us  bwd {tot_b:7.1f} us")

"""Single weight-gradient launches at benchmark size: python scripts/prof_wgrad.py  (device-timed, back-to-back)."""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from view_fusion_b200 import _lib

lib = _lib.require_device()
R = 168
cases = {"c64": (64, [(64, 3)], 64), "c128": (32, [(128, 3)], 128), "u64b": (64, [(64, 3), (128, 1)], 64), "u64a": (64, [(128, 3)], 64),
         "c192": (16, [(192, 3)], 192), "c320": (8, [(320, 3)], 320)}
for name, (S, segs, cout) in cases.items():
    rows = R * (S + 1) * (S + 1)
    srcs = [torch.randn(rows, c, device="cuda").to(torch.bfloat16) for c, _ in segs]
    dy = torch.randn(rows, cout, device="cuda").to(torch.bfloat16)
    k_total = sum(c * k * k for c, k in segs)
    a = _lib.ConvArgs()
    a.dtype, a.images, a.H, a.W, a.in_padded, a.out_padded, a.n_seg, a.stride = _lib.VF_BF16, R, S, S, 1, 1, len(segs), 1
    for i, ((c, k), s) in enumerate(zip(segs, srcs)):
        a.src[i], a.src_c[i], a.ksize[i] = s.data_ptr(), c, k
    a.cout, a.cout_pad = cout, cout
    dwp = torch.zeros(cout, k_total, device="cuda")
    flops = 2 * R * S * S * cout * k_total
    for order in (0, 1, 2):
        lib.vf_debug_flags(order << 28)
        ts = []
        for it in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            big = torch.empty(64 << 20, device="cuda").fill_(1.0)
            e0.record()
            for _ in range(10):
                _lib.check(lib.vf_conv2d_wgrad(C.byref(a), dy.data_ptr(), cout, dwp.data_ptr(), _lib.stream_handle()), "wgrad")
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 100)
        t = min(ts[1:])
        print(f"{name:5s} order {order}: {t:7.1f} us  {flops / t / 1e6:7.1f} TFLOP/s", flush=True)
lib.vf_debug_flags(0)

"""Tiling sweep of representative 3x3 layers at 168 view-images (forced (block_n, G) through vf_debug_flags)."""
import sys
import torch
sys.path.insert(0, ".")
from view_fusion_b200 import _lib, ops

lib = _lib.require_device()
R = 168
bf = torch.bfloat16
def rnd(*s): return torch.randn(*s, device="cuda").to(bf)

def timeit(fn, iters=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

layers = [("16x16 192->192 conv1", 16, [(192, 3)], 192), ("16x16 512->192 conv1", 16, [(512, 3)], 192), ("16x16 192 conv2+id", 16, [(192, 3), (192, 1)], 192),
          ("8x8 320->320 conv1", 8, [(320, 3)], 320), ("8x8 640->320 conv1", 8, [(640, 3)], 320),
          ("32x32 128->128 conv1", 32, [(128, 3)], 128), ("32x32 320->128 conv1", 32, [(320, 3)], 128), ("64x64 128->64 conv1", 64, [(128, 3)], 64)]
for name, S, segs, cout in layers:
    srcs = [rnd(R * (S + 1) * (S + 1), c) for c, _ in segs]
    kt = sum(c * k * k for c, k in segs)
    w = rnd(cout, kt) * 0.03
    out = torch.empty(R * (S + 1) * (S + 1), cout, device="cuda", dtype=bf)
    st = torch.zeros(R, cout, 2, device="cuda")
    bias = torch.randn(cout, device="cuda") if len(segs) > 1 else None          # conv2 keeps its bias; conv1's is deferred
    fn = lambda: ops.conv2d(srcs, [k for _, k in segs], w, R, S, S, cout, bias=bias, out=out, want_stats=True, stats=st)
    lib.vf_debug_flags(0)
    flops = 2.0 * R * S * S * kt * cout
    t0 = timeit(fn)
    print(f"{name:24s} model choice       : {t0:7.1f} us  {flops / t0 / 1e6:7.0f} TF")
    for bn in sorted({cout // t for t in (1, 2, 3, 4, 5) if cout % t == 0 and (cout // t) % 64 == 0}):
        for g in (1, 2, 3, 4):
            if 2 * g * bn > 512 or g * ((bn + 63) // 64) > 4:
                continue
            lib.vf_debug_flags(((bn // 16) << 20) | (g << 16))
            try:
                t = timeit(fn)
            except RuntimeError:
                continue
            print(f"{name:24s}   bn={bn:3d} G={g}        : {t:7.1f} us  {flops / t / 1e6:7.0f} TF")
lib.vf_debug_flags(0)

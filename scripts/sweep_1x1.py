"""Tiling sweep of the 1x1 layers of the benchmark UNet at 168 view-images (forced (block_n, G) through vf_debug_flags;
bit 4096 = no pinned N tile, bit 8192 = no deep activation ring).  Prints us per launch (CUDA events, 20 launches)."""
import sys
import torch
sys.path.insert(0, ".")
from view_fusion_b200 import _lib, ops

lib = _lib.require_device()
R = 168
bf = torch.bfloat16
def rnd(*s): return torch.randn(*s, device="cuda").to(bf)

def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

cases = {}
S, Cc = 16, 192
x_p = rnd(R * (S + 1) * (S + 1), Cc); w_qkv = rnd(3 * Cc, Cc) * 0.05
out_q = torch.empty(R * S * S, 3 * Cc, device="cuda", dtype=bf)
cases["qkv192"] = (lambda: ops.conv2d([x_p], [1], w_qkv, R, S, S, 3 * Cc, out_padded=False, out=out_q), [64, 192], [1, 2, 4])
o_f = rnd(R * S * S, Cc); w_o = rnd(Cc, Cc) * 0.05; b_o = torch.randn(Cc, device="cuda"); res = rnd(R * (S + 1) * (S + 1), Cc)
out_o = torch.empty(R * (S + 1) * (S + 1), Cc, device="cuda", dtype=bf); st_o = torch.zeros(R, Cc, 2, device="cuda")
cases["out192"] = (lambda: ops.conv2d([o_f], [1], w_o, R, S, S, Cc, bias=b_o, residual=res, in_padded=False, out_padded=True, out=out_o, want_stats=True, stats=st_o), [64, 192], [1, 2, 4])
x0 = rnd(R * 64 * 64, 64); w0 = rnd(64, 64) * 0.05; b0 = torch.randn(64, device="cuda")
out0 = torch.empty(R * 65 * 65, 64, device="cuda", dtype=bf); st0 = torch.zeros(R, 64, 2, device="cuda")
cases["conv0"] = (lambda: ops.conv2d([x0], [1], w0, R, 64, 64, 64, bias=b0, in_padded=False, out_padded=True, out=out0, want_stats=True, stats=st0), [64], [1, 2, 4])
for name, (fn, bns, gs) in cases.items():
    lib.vf_debug_flags(0)
    print(f"{name:8s} default(model)      : {timeit(fn):7.1f} us")
    for extra, tag in ((0, "pin+deepA"), (4096, "no pin"), (8192, "no deepA"), (4096 | 8192, "neither")):
        for bn in bns:
            for g in gs:
                if 2 * g * bn > 512 or g * ((bn + 63) // 64) > 4:
                    continue
                lib.vf_debug_flags(((bn // 16) << 20) | (g << 16) | extra)
                try:
                    t = timeit(fn)
                except RuntimeError as e:
                    t = float("nan")
                print(f"{name:8s} bn={bn:3d} G={g} {tag:10s}: {t:7.1f} us")
lib.vf_debug_flags(0)

"""GroupNorm(+Swish) backward at the benchmark's layer shapes (168 view-images, bf16): one-pass kernel vs the two-pass reduce + apply
kernels (vf_debug_flags 0x2000).  python scripts/gn_bwd_bench.py"""
import sys
import torch
sys.path.insert(0, ".")
from view_fusion_b200 import ops, _lib

BF, R, dev = torch.bfloat16, 168, "cuda"
lib = _lib.require_device()


def timed(fn, iters=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


tot = [0.0, 0.0]
for S, C0, C1, n in [(64, 64, 0, 13), (64, 128, 64, 1), (64, 64, 64, 1), (32, 64, 0, 1), (32, 128, 0, 11), (32, 192, 128, 1), (32, 128, 128, 1), (32, 128, 64, 1),
                     (16, 128, 0, 1), (16, 192, 0, 14), (16, 320, 192, 1), (16, 192, 192, 1), (16, 192, 128, 1), (8, 192, 0, 1), (8, 320, 0, 15), (8, 320, 320, 2), (8, 320, 192, 1)]:
    C, P = C0 + C1, (S + 1) * (S + 1)
    s0 = (torch.randn(R * P, C0, device=dev) * 1.5).to(BF)
    s1 = torch.randn(R * P, C1, device=dev).to(BF) if C1 else None
    dy = torch.randn(R * P, C, device=dev).to(BF)
    dx0 = torch.empty(R * P, C0, dtype=BF, device=dev)
    dx1 = torch.empty(R * P, C1, dtype=BF, device=dev) if C1 else None
    gamma, beta = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.1
    st = ops.gn_stats(s0, s1, R, S, S)
    scratch = torch.zeros(R * C * 2, device=dev)
    dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)

    def run():
        _lib.check(lib.vf_gn_backward(s0.data_ptr(), C0, st.data_ptr(), C, _lib.ptr(s1), C1, st.data_ptr() + 8 * C0 if C1 else 0, C, _lib.VF_BF16, R, S, S, 32,
                                      gamma.data_ptr(), beta.data_ptr(), 1, dy.data_ptr(), scratch.data_ptr(), dg.data_ptr(), db.data_ptr(), dx0.data_ptr(), 0,
                                      _lib.ptr(dx1), 0, None, _lib.stream_handle()), "gn_bwd")

    lib.vf_debug_flags(0x2000)
    t2 = timed(run)
    lib.vf_debug_flags(0)
    t1 = timed(run)
    mb = R * P * C * 2 / 1e6      # one tensor, MB
    tot[0] += n * t2; tot[1] += n * t1
    print(f"{S:2d}x{S:<2d} C={C0:3d}+{C1:<3d} x{n:2d}: two-pass {t2:6.1f} us | one-pass {t1:6.1f} us ({3 * mb / t1:5.2f} TB/s over 3 tensor passes)", flush=True)
print(f"per training step (layer counts of the small-v100 UNet): two-pass {tot[0] / 1e3:.2f} ms, one-pass {tot[1] / 1e3:.2f} ms")

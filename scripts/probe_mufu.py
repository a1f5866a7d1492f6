"""Special-function throughput on B200 (what bounds SiLU): python scripts/probe_mufu.py"""
import sys, torch
sys.path.insert(0, ".")
from view_fusion_b200 import _lib
_lib.require_device()
lib = _lib.load_probes()
names = ["tanh.approx.f32", "ex2 + rcp (sigmoid)", "tanh.approx.f16x2 (+cvt)", "ex2 only", "rcp only", "FMA only (loop baseline)"]
for per_sm in (1, 4, 8):
    grid = 148 * per_sm
    out = torch.zeros(grid, dtype=torch.int64, device="cuda")
    scratch = torch.zeros(grid * 256, device="cuda")
    for mode in range(6):
        iters = 2000
        for _ in range(2):
            _lib.check(lib.vf_debug_mufu_rate(mode, iters, grid, scratch.data_ptr(), out.data_ptr(), _lib.stream_handle()), "mufu")
        torch.cuda.synchronize()
        cyc = out.double().mean().item()
        el = iters * 8 * 256 * per_sm
        print(f"{per_sm} CTAs/SM  {names[mode]:28s}: {el / cyc:6.2f} elements / clk / SM", flush=True)

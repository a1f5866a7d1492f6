import sys, torch
sys.path.insert(0, ".")
from view_fusion_b200 import _lib
lib = _lib.require_device()
out = torch.zeros(148, dtype=torch.int64, device="cuda")
for N in (64, 128, 192, 256):
    for ce in (0, -2, -100, -101):
        ng = 1600
        _lib.check(_lib.load_probes().vf_debug_umma_rate(N, 3, ng, ce, 148, out.data_ptr(), _lib.stream_handle()), "rate")
        torch.cuda.synchronize()
        cyc = out.double().mean().item()
        print(f"N {N:3d} mode {ce:4d}: {cyc / (ng * 4):7.1f} cycles/MMA", flush=True)

print("weight-gradient pattern (MN-major operands, 5 accumulators):")
for N in (64, 96):
    for ce in (-200, -201, -202, -203, -204):
        ng = 1600
        _lib.check(_lib.load_probes().vf_debug_umma_rate(N, 3, ng, ce, 148, out.data_ptr(), _lib.stream_handle()), "rate")
        torch.cuda.synchronize()
        cyc = out.double().mean().item()
        print(f"N {N:3d} mode {ce:4d}: {cyc / (ng * 4):7.1f} cycles/MMA", flush=True)

print("weight-stationary pattern (-300 plain G=2, -301 .ws G=2, -302 .ws fill/lastuse G=2, -303 plain G=4, -304 .ws G=4, -305 .ws fill/use/lastuse G=4):")
for N in (64, 128):
    for ce in (-300, -301, -302, -303, -304, -305):
        ng = 1600
        _lib.check(_lib.load_probes().vf_debug_umma_rate(N, 3, ng, ce, 148, out.data_ptr(), _lib.stream_handle()), "rate")
        torch.cuda.synchronize()
        cyc = out.double().mean().item()
        print(f"N {N:3d} mode {ce:4d}: {cyc / (ng * 4):7.1f} cycles/MMA", flush=True)

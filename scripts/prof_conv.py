"""Single conv launches at benchmark size: python scripts/prof_conv.py case [ablate]  (device-timed, pre-allocated)."""
import math
import sys
import torch
sys.path.insert(0, ".")
from view_fusion_b200 import ops, _lib

torch.manual_seed(0)
R = 168
case = sys.argv[1] if len(sys.argv) > 1 else "c64"
cfgs = {
    "c0": dict(S=64, segs=[(64, 1)], cout=64, flat=True),
    "c64": dict(S=64, segs=[(64, 3)], cout=64, flat=False),
    "c128": dict(S=32, segs=[(128, 3)], cout=128, flat=False),
    "c192": dict(S=16, segs=[(192, 3)], cout=192, flat=False),
    "c320": dict(S=8, segs=[(320, 3)], cout=320, flat=False),
    "u8a": dict(S=8, segs=[(640, 3)], cout=320, flat=False),
    "u8b": dict(S=8, segs=[(320, 3), (640, 1)], cout=320, flat=False),
    "u16a": dict(S=16, segs=[(512, 3)], cout=192, flat=False),
    "u32a": dict(S=32, segs=[(320, 3)], cout=128, flat=False),
    "u64": dict(S=64, segs=[(64, 3), (64, 1), (64, 1)], cout=64, flat=False),
    "d64": dict(S=64, segs=[(64, 3)], cout=64, flat=False, stride=2),
    "d128": dict(S=32, segs=[(128, 3)], cout=128, flat=False, stride=2),
    "d192": dict(S=16, segs=[(192, 3)], cout=192, flat=False, stride=2),
}
c = cfgs[case]
S, segs, cout = c["S"], c["segs"], c["cout"]
stride = c.get("stride", 1)
So = S // stride
rows = R * (S * S if c["flat"] else (S + 1) * (S + 1))
srcs = [torch.randn(rows, ch, device="cuda").to(torch.bfloat16) for ch, _ in segs]
k_total = sum(ch * k * k for ch, k in segs)
w = (torch.randn(cout, k_total, device="cuda") / math.sqrt(k_total)).to(torch.bfloat16)
bias = torch.randn(cout, device="cuda")
res = torch.randn(R * (So + 1) * (So + 1), cout, device="cuda").to(torch.bfloat16)
out = torch.empty(R * (So + 1) * (So + 1), cout, device="cuda", dtype=torch.bfloat16)
stats = torch.zeros(R, cout, 2, device="cuda")
flops = 2 * R * So * So * cout * k_total
variants = [("full", 0, True, True)]
if len(sys.argv) > 2 and sys.argv[2] == "ablate":
    variants += [("no-res (typical)", 0, False, True), ("no-res dbg:no-table", 8, False, True), ("no-res-no-stats", 0, False, False), ("dbg:no-store", 2, True, True), ("dbg:no-unit-work", 4, True, True), ("dbg:no-unit-no-table", 12, True, True)]
if len(sys.argv) > 2 and sys.argv[2] == "r02":
    variants = [("conv1: no table", 0, False, True, False), ("bias only", 0, False, True, True), ("conv2: bias+residual", 0, True, True, True),
                ("no table, no stats", 0, False, False, False)]
variants = [v if len(v) == 5 else (*v, True) for v in variants]
if len(sys.argv) > 2 and sys.argv[2] == "tail":
    variants = [("short last box", 0, False, True, True), ("64-row boxes", 0x4000, False, True, True)]
if len(sys.argv) > 2 and sys.argv[2] == "plain":
    variants = [("bias + stats", 0, False, True, True)]
if len(sys.argv) > 2 and sys.argv[2] == "typical":
    variants = [("no-res (typical)", 0, False, True)]
if len(sys.argv) > 2 and sys.argv[2] == "sweep":
    variants = []
    for bn in sorted({b for b in (64, 96, 128, 160, 192, 256, 320) if cout % b == 0 and b <= 256}):
        for G in (1, 2, 3, 4):
            for noresid in (0, 512):
                variants.append((f"bn{bn} G{G} {'ring' if noresid else 'resident-ok'}", (bn // 16) << 20 | G << 16 | noresid | 256, True, True))
lib = _lib.require_device()
variants = [v if len(v) == 5 else (*v, True) for v in variants]
for name, dbg, use_res, use_stats, use_bias in variants:
    lib.vf_debug_flags(dbg)
    try:
        ops.conv2d(srcs, [k for _, k in segs], w, R, S, S, cout, bias=bias if use_bias else None, residual=res if use_res else None, want_stats=use_stats,
                   in_padded=not c["flat"], out=out, stats=stats, stride=stride)
        torch.cuda.synchronize()
    except RuntimeError as e:
        continue
    lib.vf_debug_flags(dbg & ~256)
    ts = []
    NL = 20
    for it in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        big = torch.empty(64 << 20, device="cuda").fill_(1.0)      # keeps the GPU busy while the launches are enqueued
        e0.record()
        for _ in range(NL):
            ops.conv2d(srcs, [k for _, k in segs], w, R, S, S, cout, bias=bias if use_bias else None, residual=res if use_res else None, want_stats=use_stats,
                       in_padded=not c["flat"], out=out, stats=stats, stride=stride)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / NL)
    t = min(ts[1:])
    cnt = torch.zeros(148 * 12 + 1 + 4 * 64, dtype=torch.int64, device="cuda")
    lib.vf_debug_counters(cnt.data_ptr())
    for _ in range(6):
        ops.conv2d(srcs, [k for _, k in segs], w, R, S, S, cout, bias=bias if use_bias else None, residual=res if use_res else None, want_stats=use_stats,
                   in_padded=not c["flat"], out=out, stats=stats, stride=stride)
    torch.cuda.synchronize()
    lib.vf_debug_counters(0)
    st = cnt[148 * 12 + 1: 148 * 12 + 1 + 24].view(6, 4).cpu().double()
    setup = (st[1:, 1] - st[1:, 0]).mean().item() / 1e3; main = (st[1:, 2] - st[1:, 1]).mean().item() / 1e3
    tear = (st[1:, 3] - st[1:, 2]).mean().item() / 1e3; gap = (st[1:, 0] - st[:-1, 3]).mean().item() / 1e3
    period = (st[1:, 0] - st[:-1, 0]).mean().item() / 1e3
    stamp_txt = f" | cta0 us: entry->ready {setup:5.1f} loops {main:5.1f} teardown {tear:4.1f} exit->next entry {gap:5.1f} period {period:5.1f}"
    cm = cnt[:148 * 4].view(148, 4).double().mean(0).tolist()
    em = cnt[148 * 4:148 * 12].view(148, 8).double().mean(0).tolist()
    print(f"{case:5s} {name:22s}: {t:7.1f} us  {flops / t / 1e6:7.1f} TFLOP/s | MMA thread kcyc: total {cm[0]/1e3:6.1f} waitA {cm[1]/1e3:6.1f} waitB {cm[2]/1e3:6.1f} waitAcc {cm[3]/1e3:6.1f} | epi w0 kcyc: total {em[0]/1e3:6.1f} table {em[1]/1e3:5.1f} waitAcc {em[2]/1e3:6.1f} ld+pack {em[3]/1e3:5.1f} (store-wait {em[4]/1e3:5.1f}) stats {em[5]/1e3:5.1f} prep {em[6]/1e3:4.1f}{stamp_txt}")
lib.vf_debug_flags(0)

"""GPU: training forward/backward through the drop-in ViewFusion module vs the reference goldens
(loss, per-parameter gradient norms, sampled gradient values; produced by the reference's autograd)."""
import pytest
import torch

import vf_oracle as O
from gpu_util import build_model, load, rel

pytestmark = pytest.mark.gpu


def _run(golden_dir, tag, cfg, prec):
    g = load(golden_dir, f"train_{tag}")
    m, _ = build_model(cfg, int(g["seed"]), prec)
    loss = m(y_cond=g["y_cond"].cuda(), view_count=g["view_count"], angle=g["angle"].cuda(), y_0=g["y_0"].cuda(), noise=g["noise"].cuda(),
             t=g["t"].cuda(), u=g["u"].cuda())
    loss.backward()
    torch.cuda.synchronize()
    grads = {n: p.grad.detach().cpu() for n, p in m.denoise_fn.named_parameters()}
    return g, float(loss.detach()), grads


@pytest.mark.parametrize("tag,cfg", [("tiny_ragged", O.TINY), ("small_n3", O.SMALL_V100)])
def test_train_fp32_mode_loss_and_gradients(golden_dir, tag, cfg):
    g, loss, grads = _run(golden_dir, tag, cfg, "fp32")
    assert abs(loss - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))
    names = [str(n) for n in g["grad_names"]]
    norms = g["grad_norms"].tolist()
    big = max(norms)
    bad = {}
    for n, ref in zip(names, norms):
        mine = float(grads[n].norm())
        if ref < 1e-6 * big:
            if not mine < 1e-5 * big:
                bad[n] = (mine, ref)
        elif not abs(mine - ref) < 2e-3 * ref:
            bad[n] = (mine, ref)
    assert not bad, f"{len(bad)} gradient norms off, e.g. {list(bad.items())[:6]}"
    for k in g:
        if k.startswith("grad:"):
            full = grads[k[5:]].reshape(-1)
            stride = max(1, (full.numel() + 8191) // 8192)
            assert rel(full[::stride], g[k]) < 2e-3, k


def test_train_bf16_mode_loss_and_gradient_direction(golden_dir):
    """bf16 tensor-core training step vs the reference's fp32 autograd (golden).  Per-tensor bars: every gradient tensor that
    carries signal (norm > 1e-3 of the largest) within 3 % in norm (zero exceptions) and cos > 0.999 of the stored sample; the measured worst
    cases go to the margins file (the kernels are pinned one by one, at tight bars, in tests/test_gpu_bwd_ops.py)."""
    from gpu_util import margin
    g, loss, grads = _run(golden_dir, "small_n3", O.SMALL_V100, "bf16")
    ok = margin("train small-v100 N=3 bf16: loss relative error vs reference", abs(loss - float(g["loss"])) / abs(float(g["loss"])), 1e-2)
    names = [str(n) for n in g["grad_names"]]
    norms = dict(zip(names, g["grad_norms"].tolist()))
    big = max(norms.values())
    live = {n: r for n, r in norms.items() if r > 1e-3 * big}
    dev = {n: abs(float(grads[n].norm()) - r) / r for n, r in live.items()}
    worst = max(dev, key=dev.get)
    ok &= margin(f"train small-v100 N=3 bf16: worst per-tensor gradient-norm deviation ({worst})", dev[worst], 3e-2)
    ok &= margin("train small-v100 N=3 bf16: mean per-tensor gradient-norm deviation", sum(dev.values()) / len(dev), 1e-2)
    coss = {}
    for k in g:
        if k.startswith("grad:"):
            full = grads[k[5:]].reshape(-1)
            stride = max(1, (full.numel() + 8191) // 8192)
            a, b = full[::stride].double(), g[k].double()
            if float(b.norm()) > 1e-3 * big:
                coss[k[5:]] = float((a * b).sum() / (a.norm() * b.norm()))
    wc = min(coss, key=coss.get)
    ok &= margin(f"train small-v100 N=3 bf16: worst gradient cosine vs reference ({wc})", coss[wc], 0.999, higher_is_better=True)
    assert ok


@pytest.mark.parametrize("case", [(3, 16, [(64, 3)], 64), (2, 32, [(128, 3)], 128), (3, 16, [(128, 3), (64, 1), (192, 1)], 192),
                                  (5, 8, [(320, 3)], 320), (2, 64, [(64, 3), (64, 1)], 64), (3, 32, [(128, 3), (64, 1), (64, 1)], 64),
                                  (40, 16, [(192, 3)], 64)])        # Cout <= 64: the kernel with the three kernel rows in N
def test_wgrad_tcgen05_vs_cuda_cores(case):
    """tcgen05 MN-major weight-gradient GEMM vs the CUDA-core kernel and vs fp64 math on the same bf16 operands."""
    import ctypes as C
    from view_fusion_b200 import _lib, ops
    lib = _lib.require_device()
    R, S, segs, cout = case
    torch.manual_seed(1)
    bf = lambda t: t.to(torch.bfloat16)
    xs = [bf(torch.randn(R, c, S, S)) for c, _ in segs]
    dy = bf(torch.randn(R, cout, S, S))
    srcs = [ops.to_padded(x.float(), torch.bfloat16).cuda() for x in xs]       # zero padding rows
    dyp = ops.to_padded(dy.float(), torch.bfloat16).cuda()
    k_total = sum(c * k * k for c, k in segs)
    a = _lib.ConvArgs()
    a.dtype, a.images, a.H, a.W, a.in_padded, a.out_padded, a.n_seg, a.stride = _lib.VF_BF16, R, S, S, 1, 1, len(segs), 1
    for i, ((c, k), s) in enumerate(zip(segs, srcs)):
        a.src[i], a.src_c[i], a.ksize[i] = s.data_ptr(), c, k
    a.cout, a.cout_pad = cout, cout
    outs = []
    for simt in (False, True):
        ops.force_simt(simt)
        try:
            dwp = torch.zeros(cout, k_total, device="cuda")
            _lib.check(lib.vf_conv2d_wgrad(C.byref(a), dyp.data_ptr(), cout, dwp.data_ptr(), _lib.stream_handle()), "wgrad")
            torch.cuda.synchronize()
            outs.append(dwp.cpu())
        finally:
            ops.force_simt(False)
    # reference: dW[n][tap][c] = sum_pixels dy[n] * x[c] shifted
    ref = torch.zeros(cout, k_total, dtype=torch.float64)
    off = 0
    for (c, k), x in zip(segs, xs):
        xd = torch.nn.functional.unfold(x.double(), k, padding=k // 2).view(R, c, k * k, S * S)          # (R, c, taps, L)
        g = torch.einsum("rnl,rctl->ntc", dy.double().view(R, cout, S * S), xd).reshape(cout, k * k * c)
        ref[:, off:off + k * k * c] = g
        off += k * k * c
    assert rel(outs[1], ref.float()) < 1e-4, "CUDA-core weight gradient"
    assert rel(outs[0], ref.float()) < 1e-4, "tcgen05 weight gradient"


@pytest.mark.parametrize("R,L,Cc", [(3, 256, 192), (2, 128, 128), (5, 256, 128), (2, 64, 128), (3, 64, 320)])
def test_attention_backward_bf16(R, L, Cc):
    """Attention backward (tcgen05 where the shape fits, CUDA cores otherwise) vs autograd on the same bf16 operands."""
    import math
    from gpu_util import bf16r
    from view_fusion_b200 import ops
    torch.manual_seed(5)
    qkv = bf16r(torch.randn(R, L, 3 * Cc) * 1.2).double().requires_grad_(True)
    d_out = bf16r(torch.randn(R, L, Cc))
    q, k, v = qkv[..., :Cc], qkv[..., Cc:2 * Cc], qkv[..., 2 * Cc:]
    o = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(Cc), dim=-1) @ v
    o.backward(d_out.double())
    ref = qkv.grad.reshape(R * L, 3 * Cc).float()
    qd = qkv.detach().reshape(R * L, 3 * Cc).to(torch.bfloat16).cuda()
    vt = v.detach().transpose(1, 2).contiguous().to(torch.bfloat16).cuda()
    lse = torch.zeros(R * L, device="cuda")
    out = ops.attention(qd, vt, R, L, Cc, lse=lse)
    dod = d_out.reshape(R * L, Cc).to(torch.bfloat16).cuda()
    got = ops.attention_backward(qd, vt, out, lse, dod, R, L, Cc)
    ops.force_simt(True)
    try:
        simt = ops.attention_backward(qd, vt, None, None, dod, R, L, Cc)
    finally:
        ops.force_simt(False)
    torch.cuda.synchronize()
    for name, sl in (("dq", slice(0, Cc)), ("dk", slice(Cc, 2 * Cc)), ("dv", slice(2 * Cc, 3 * Cc))):
        assert rel(simt[:, sl], ref[:, sl]) < 1e-2, ("cuda cores", name)
        assert rel(got[:, sl], ref[:, sl]) < 1.5e-2, ("tcgen05", name)


@pytest.mark.parametrize("case", [(3, 16, 192, 576, 576, False), (2, 64, 64, 64, 64, False), (2, 32, 128, 6, 64, True)])
def test_wgrad_tcgen05_1x1_flat_rows_and_padded_columns(case):
    """1x1 weight gradient on FLAT rows (qkv / out-projection / first layer) and with dY columns beyond cout (final conv)."""
    import ctypes as C
    from view_fusion_b200 import _lib, ops
    lib = _lib.require_device()
    R, S, cin, cout, ld, padded = case
    torch.manual_seed(2)
    x = torch.randn(R, cin, S, S).to(torch.bfloat16)
    dy = torch.zeros(R, ld, S, S)
    dy[:, :cout] = torch.randn(R, cout, S, S)
    dy = dy.to(torch.bfloat16)
    if padded:
        xs, dys = ops.to_padded(x.float(), torch.bfloat16).cuda(), ops.to_padded(dy.float(), torch.bfloat16).cuda()
    else:
        xs = x.permute(0, 2, 3, 1).reshape(-1, cin).contiguous().cuda()
        dys = dy.permute(0, 2, 3, 1).reshape(-1, ld).contiguous().cuda()
    a = _lib.ConvArgs()
    a.dtype, a.images, a.H, a.W, a.in_padded, a.out_padded, a.n_seg, a.stride = _lib.VF_BF16, R, S, S, int(padded), int(padded), 1, 1
    a.src[0], a.src_c[0], a.ksize[0] = xs.data_ptr(), cin, 1
    a.cout, a.cout_pad = cout, ld
    dwp = torch.zeros(ld, cin, device="cuda")
    _lib.check(lib.vf_conv2d_wgrad(C.byref(a), dys.data_ptr(), ld, dwp.data_ptr(), _lib.stream_handle()), "wgrad")
    torch.cuda.synchronize()
    ref = torch.einsum("rnl,rcl->nc", dy.double().view(R, ld, S * S), x.double().view(R, cin, S * S)).float()
    assert rel(dwp.cpu()[:cout], ref[:cout]) < 1e-4
    assert float(dwp[cout:].abs().max()) == 0.0 if cout < ld else True


def test_fused_adam_matches_torch_adam():
    """view_fusion_b200.optim.FusedAdam vs torch.optim.Adam over several steps (odd sizes, unaligned views, weight decay)."""
    from view_fusion_b200.optim import FusedAdam
    torch.manual_seed(0)
    shapes = [(64, 6, 3, 3), (64,), (320, 640, 1, 1), (5568, 64), (7,), (4099,)]
    flat = torch.randn(sum(int(torch.tensor(s).prod()) for s in shapes) + 3, device="cuda")
    for wd in (0.0, 0.01):
        ref_p = [torch.nn.Parameter(torch.randn(*s, device="cuda")) for s in shapes]
        my_p = [torch.nn.Parameter(p.detach().clone()) for p in ref_p]
        ref = torch.optim.Adam(ref_p, lr=3e-3, weight_decay=wd)
        mine = FusedAdam(my_p, lr=3e-3, weight_decay=wd)
        for it in range(5):
            off = 3                                   # gradients are (possibly misaligned) views of one flat buffer
            for a, b in zip(ref_p, my_p):
                g = torch.randn_like(a)
                a.grad = g.clone()
                view = flat[off:off + g.numel()].view_as(g)
                view.copy_(g)
                b.grad = view
                off += g.numel()
            if it == 3:
                for o in (ref, mine):
                    o.param_groups[0]["lr"] = 1e-3    # scheduler-style learning-rate rewrite
            ref.step()
            mine.step()
        for a, b in zip(ref_p, my_p):
            assert rel(b, a) < 2e-6
        sd = mine.state_dict()
        assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"} and float(sd["state"][0]["step"]) == 5.0
        assert rel(mine.state[my_p[2]]["exp_avg_sq"], ref.state[ref_p[2]]["exp_avg_sq"]) < 2e-6

"""GPU: stage-level operators through the C ABI against the CPU oracle / golden fixtures."""
import ctypes as C
import math

import pytest
import torch
import torch.nn.functional as F

import vf_oracle as O
from gpu_util import bf16r, load, rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from view_fusion_b200 import _lib
    return _lib.require_device()


def _compose(lib, out_nchw, view_count, y_t, t, z, weighting, add_noise=True, want=True):
    from view_fusion_b200 import _lib
    dev = "cuda"
    R, Cc, H, W = out_nchw.shape
    o8 = torch.zeros(R, H, W, 8)
    o8[..., :Cc] = out_nchw.permute(0, 2, 3, 1)
    o8 = o8.to(dev).contiguous()
    B = view_count.numel()
    off = torch.zeros(B + 1, dtype=torch.int32)
    off[1:] = torch.cumsum(view_count, 0)
    sched = {k: v.to(dev) for k, v in O.make_schedule(**O.BETA_TRAIN).items()}
    s = _lib.Schedule()
    for k, v in sched.items():
        setattr(s, k, v.data_ptr())
    s.num_timesteps = 2000
    a = _lib.ComposeArgs()
    offd, t32, ytd = off.to(dev), t.to(dev, torch.int32), y_t.to(dev).contiguous()
    y_prev = torch.empty_like(ytd)
    eps = torch.empty_like(ytd)
    mv = int(view_count.max())
    w = torch.full((B, mv, 3, H, W), -7.0, device=dev) if (weighting and want) else None
    lg = torch.empty(R, 3, H, W, device=dev) if (weighting and want) else None
    zd = None if z is None else z.to(dev).contiguous()
    a.unet_out, a.view_offset, a.t, a.y_t, a.y_prev, a.z = o8.data_ptr(), offd.data_ptr(), t32.data_ptr(), ytd.data_ptr(), y_prev.data_ptr(), _lib.ptr(zd)
    a.seed, a.offset = 1234, 1
    a.add_noise, a.clip_denoised, a.weighting, a.B, a.H, a.W = int(add_noise), 1, int(weighting), B, H, W
    a.eps_out, a.weights_out, a.max_v, a.logits_out = eps.data_ptr(), _lib.ptr(w), mv, _lib.ptr(lg)
    _lib.check(lib.vf_compose_ddpm_step(C.byref(a), C.byref(s), _lib.stream_handle()), "compose")
    torch.cuda.synchronize()
    return y_prev.cpu(), eps.cpu(), (None if w is None else w.cpu()), (None if lg is None else lg.cpu())


@pytest.mark.parametrize("tag,weighting", [("ragged", True), ("full6", True), ("mean", False)])
def test_compose_ddpm_against_reference_golden(lib, golden_dir, tag, weighting):
    g = load(golden_dir, f"compose_{tag}")
    for tn in ("hi", "mid", "one", "zero"):
        t = g[f"t_{tn}"]
        y_prev, eps, w, lg = _compose(lib, g["out"], g["view_count"], g["y_t"], t, g["z"], weighting, add_noise=bool((t > 0).any()))
        assert rel(eps, g["eps"]) < 2e-6
        assert rel(y_prev, g[f"y_prev_{tn}"]) < 2e-6, tn
        if weighting:
            assert rel(w, g["weights"]) < 2e-6
            assert torch.equal(lg, g["out"][:, 3:])
            assert float(w.min()) >= 0.0            # padded slots were overwritten with exact zeros


def test_compose_philox_noise_is_standard_normal(lib):
    B, V, S = 4, 3, 64
    out = torch.zeros(B * V, 6, S, S)
    y_t = torch.zeros(B, 3, S, S)
    vc = torch.full((B,), V)
    t = torch.full((B,), 1000)
    y_prev, _, _, _ = _compose(lib, out, vc, y_t, t, None, True, want=False)
    sched = O.make_schedule(**O.BETA_TRAIN)
    z = y_prev / math.exp(0.5 * float(sched["posterior_log_variance_clipped"][1000]))
    assert abs(float(z.mean())) < 0.02 and abs(float(z.std()) - 1.0) < 0.02
    assert abs(float((z ** 4).mean()) - 3.0) < 0.15
    c = torch.corrcoef(torch.stack([z[:, 0].flatten(), z[:, 1].flatten(), z[:, 2].flatten()]))
    assert float((c - torch.eye(3)).abs().max()) < 0.02


def test_compose_mse_and_grad(lib):
    from view_fusion_b200 import _lib
    torch.manual_seed(0)
    vc = torch.tensor([2, 5, 1, 6])
    B, R, S = 4, int(vc.sum()), 16
    out = torch.randn(R, 6, S, S, requires_grad=True)
    noise = torch.randn(B, 3, S, S)
    eps, _, _ = O.compose(out, vc, True)
    loss = F.mse_loss(noise, eps)
    loss.backward()
    o8 = torch.zeros(R, S, S, 8)
    o8[..., :6] = out.detach().permute(0, 2, 3, 1)
    o8 = o8.cuda()
    off = torch.zeros(B + 1, dtype=torch.int32)
    off[1:] = torch.cumsum(vc, 0)
    offd, nd = off.cuda(), noise.cuda()
    lacc = torch.zeros(1, device="cuda")
    epsd = torch.empty(B, 3, S, S, device="cuda")
    grad = torch.empty_like(o8)
    _lib.check(lib.vf_compose_mse(o8.data_ptr(), offd.data_ptr(), nd.data_ptr(), B, S, S, 1, lacc.data_ptr(), epsd.data_ptr(),
                                  grad.data_ptr(), 1.0, _lib.stream_handle()), "mse")
    assert abs(float(lacc) - float(loss)) < 1e-6 * float(loss) + 1e-7
    assert rel(epsd, eps) < 2e-6
    g = grad.cpu()[..., :6].permute(0, 3, 1, 2)
    assert rel(g, out.grad) < 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("C0,C1,HW,groups", [(64, 0, 256, 32), (192, 128, 64, 32), (128, 64, 1024, 32), (320, 320, 64, 32), (32, 0, 256, 32)])
def test_groupnorm_swish(lib, dtype, C0, C1, HW, groups):
    from view_fusion_b200 import ops
    if dtype == torch.bfloat16 and (C0 % 8 or C1 % 8):
        pytest.skip("vector width")
    torch.manual_seed(1)
    R = 3
    S = int(math.isqrt(HW))
    x0 = torch.randn(R, C0, S, S) * 1.5 + 0.3
    x1 = torch.randn(R, C1, S, S) * 0.7 - 0.2 if C1 else None
    gamma, beta = torch.rand(C0 + C1) + 0.5, torch.randn(C0 + C1) * 0.1
    if dtype == torch.bfloat16:
        x0, x1 = bf16r(x0), (None if x1 is None else bf16r(x1))
    xc = x0 if x1 is None else torch.cat([x0, x1], 1)
    for swish in (True, False):
        ref = F.group_norm(xc, groups, gamma, beta, eps=1e-5)
        ref = O.swish(ref) if swish else ref
        # padding rows of the sources hold garbage (conv outputs never write them): they must not be read
        s0 = ops.to_padded(x0, dtype, fill=float("nan")).cuda()
        s1 = None if x1 is None else ops.to_padded(x1, dtype, fill=float("nan")).cuda()
        st = ops.gn_stats(s0, s1, R, S, S)
        y = ops.gn_apply(s0, s1, R, S, S, groups, st, gamma.cuda(), beta.cuda(), swish)
        got = ops.from_padded(y, R, S, S)
        assert rel(got, ref) < (4e-3 if dtype == torch.bfloat16 else 2e-6)
        pad = y.float().view(R, S + 1, S + 1, -1)
        assert float(pad[:, 0].abs().max()) == 0.0 and float(pad[:, :, 0].abs().max()) == 0.0   # exact zero padding rows


def _conv_case(dtype, R, S, segs, cout, stride, use_emb, use_res, seed=0):
    """segs: list of (C, ksize).  Returns (cpu reference NCHW, kwargs for ops.conv2d)."""
    torch.manual_seed(seed)
    rnd = (lambda *s: bf16r(torch.randn(*s))) if dtype == torch.bfloat16 else (lambda *s: torch.randn(*s))
    Sin = S * stride
    assert stride == 1 or len(segs) == 1
    xs = [rnd(R, c, Sin, Sin) for i, (c, k) in enumerate(segs)]
    ws = [rnd(cout, c, k, k) / math.sqrt(c * k * k) for (c, k) in segs]
    if dtype == torch.bfloat16:
        ws = [bf16r(w) for w in ws]
    bias = torch.randn(cout)
    ref = bias.view(1, -1, 1, 1).expand(R, cout, S, S).clone()
    for i, ((c, k), x, w) in enumerate(zip(segs, xs, ws)):
        ref = ref + F.conv2d(x, w, stride=stride if i == 0 else 1, padding=k // 2)
    emb = img_row = res = None
    if use_emb:
        emb = torch.randn(2, cout + 5)
        img_row = torch.randint(0, 2, (R,), dtype=torch.int32)
        ref = ref + emb[img_row.long(), :cout].view(R, cout, 1, 1)
    if use_res:
        res = rnd(R, cout, S, S)
        ref = ref + res
    return ref, xs, ws, bias, emb, img_row, res


def _run_conv(dtype, R, S, segs, cout, stride, xs, ws, bias, emb, img_row, res, **kw):
    from view_fusion_b200 import ops
    k_total = sum(c * k * k for c, k in segs)
    wp = torch.zeros(kw.get("cout_pad", cout), k_total, dtype=dtype, device="cuda")
    off = 0
    for (c, k), w in zip(segs, ws):
        ops.pack_conv_weight(w.cuda(), dtype, cout_pad=wp.shape[0], k_total=k_total, k_off=off, dst=wp)
        off += c * k * k
    Sin = S * stride
    srcs = []
    for i, ((c, k), x) in enumerate(zip(segs, xs)):
        # 3x3 sources need real zeros in the padding rows; 1x1 sources may hold anything there
        srcs.append(ops.to_padded(x, dtype, fill=0.0 if k == 3 else 7.0).cuda())
    out = ops.conv2d(srcs, [k for _, k in segs], wp, R, Sin, Sin, cout, stride=stride,
                     bias=bias.cuda(), emb=None if emb is None else emb.cuda(), img_row=None if img_row is None else img_row.cuda(),
                     residual=None if res is None else ops.to_padded(res, dtype, fill=float("nan")).cuda(), **kw)
    torch.cuda.synchronize()
    return out


CONV_CASES = [
    # R, S, segs, cout, stride, emb, res
    (2, 16, [(64, 3)], 64, 1, True, False),
    (3, 8, [(128, 3)], 128, 1, False, True),
    (2, 16, [(64, 3), (128, 1), (64, 1)], 64, 1, False, False),
    (2, 8, [(64, 3)], 64, 2, False, False),
    (1, 32, [(64, 1)], 192, 1, False, False),
    (5, 4, [(320, 3), (320, 1), (320, 1)], 320, 1, True, False),
    (1, 64, [(64, 3)], 64, 1, True, True),
    (6, 64, [(64, 3)], 64, 1, True, True),          # many work items per CTA, G=4 accumulators, image-straddling warps
    (40, 16, [(192, 3), (128, 1)], 192, 1, True, False),
    (150, 8, [(320, 3)], 320, 1, False, True),      # more items than SMs at the smallest resolution
    (3, 32, [(128, 3)], 128, 2, False, False),
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_fp32_cuda_core(lib, case):
    from view_fusion_b200 import ops
    R, S, segs, cout, stride, ue, ur = case
    ref, *t = _conv_case(torch.float32, R, S, segs, cout, stride, ue, ur)
    out = _run_conv(torch.float32, R, S, segs, cout, stride, *t)
    assert rel(ops.from_padded(out, R, S, S), ref) < 5e-6


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_bf16_tcgen05(lib, case):
    """tcgen05 kernel vs (a) exact fp32 math on the same bf16-rounded operands, (b) the CUDA-core kernel."""
    from view_fusion_b200 import ops
    R, S, segs, cout, stride, ue, ur = case
    ref, *t = _conv_case(torch.bfloat16, R, S, segs, cout, stride, ue, ur)
    out = _run_conv(torch.bfloat16, R, S, segs, cout, stride, *t)
    got = ops.from_padded(out, R, S, S)
    assert rel(got, ref) < 4e-3, "tensor-core path vs fp32 math on bf16 operands (output rounding only)"
    ops.force_simt(True)
    try:
        simt = ops.from_padded(_run_conv(torch.bfloat16, R, S, segs, cout, stride, *t), R, S, S)
    finally:
        ops.force_simt(False)
    assert rel(got, simt) < 3e-3


@pytest.mark.parametrize("case", [CONV_CASES[0], CONV_CASES[2], CONV_CASES[7], CONV_CASES[8]])
def test_conv_bf16_short_last_activation_box_is_bitwise_neutral(lib, case):
    """The last TMA box of an activation slab is only as tall as the slab needs (a second tensor map per segment); with
    vf_debug_flags(0x4000) every box is 64 rows as in round 1.  Same operands, same K order: identical bits."""
    R, S, segs, cout, stride, ue, ur = case
    _, *t = _conv_case(torch.bfloat16, R, S, segs, cout, stride, ue, ur)
    a = _run_conv(torch.bfloat16, R, S, segs, cout, stride, *t)
    lib.vf_debug_flags(0x4000)
    try:
        b = _run_conv(torch.bfloat16, R, S, segs, cout, stride, *t)
    finally:
        lib.vf_debug_flags(0)
    from view_fusion_b200 import ops
    assert torch.equal(ops.from_padded(a, R, S, S), ops.from_padded(b, R, S, S))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_conv_flat_source_padded_output(lib, dtype):
    """1x1 conv over a FLAT source (attention output / packed first layer) into a PADDED output + PADDED residual."""
    from view_fusion_b200 import ops
    torch.manual_seed(5)
    R, S, Cc, cout = 3, 16, 128, 128
    rnd = (lambda *s: bf16r(torch.randn(*s))) if dtype == torch.bfloat16 else (lambda *s: torch.randn(*s))
    x, w, res, bias = rnd(R, Cc, S, S), rnd(cout, Cc, 1, 1) / math.sqrt(Cc), rnd(R, cout, S, S), torch.randn(cout)
    ref = F.conv2d(x, w, bias) + res
    wp = ops.pack_conv_weight(w.cuda(), dtype)
    out = ops.conv2d([ops.to_nhwc(x, dtype).cuda()], [1], wp, R, S, S, cout, bias=bias.cuda(),
                     residual=ops.to_padded(res, dtype, fill=float("nan")).cuda(), in_padded=False, out_padded=True)
    assert rel(ops.from_padded(out, R, S, S), ref) < (4e-3 if dtype == torch.bfloat16 else 5e-6)


@pytest.mark.parametrize("R,S,Cc,cout", [(5, 16, 192, 192), (7, 8, 320, 320), (150, 16, 64, 128), (3, 32, 128, 64)])
def test_conv_bf16_flat_to_padded_staged_epilogue(lib, R, S, Cc, cout):
    """FLAT rows -> PADDED output through the staged TMA epilogue (32-row tiles = whole image lines of 3D tensor maps):
    output, residual fetch and fused GroupNorm sums vs exact math; padding rows of the output stay untouched."""
    from view_fusion_b200 import ops
    torch.manual_seed(11)
    dtype = torch.bfloat16
    rnd = lambda *s: bf16r(torch.randn(*s))
    x, w, res, bias = rnd(R, Cc, S, S), bf16r(rnd(cout, Cc, 1, 1) / math.sqrt(Cc)), rnd(R, cout, S, S), torch.randn(cout)
    ref = F.conv2d(x, w, bias) + res
    wp = ops.pack_conv_weight(w.cuda(), dtype)
    out, stats = ops.conv2d([ops.to_nhwc(x, dtype).cuda()], [1], wp, R, S, S, cout, bias=bias.cuda(),
                            residual=ops.to_padded(res, dtype, fill=float("nan")).cuda(), in_padded=False, out_padded=True,
                            want_stats=True)
    got = ops.from_padded(out, R, S, S).cpu()
    assert rel(got, ref) < 4e-3
    want = torch.stack([got.sum(dim=(2, 3)), (got * got).sum(dim=(2, 3))], dim=-1)
    assert rel(stats, want) < 1e-5
    pad = out.float().view(R, S + 1, S + 1, cout)
    assert torch.isnan(pad[:, 0]).all() and torch.isnan(pad[:, :, 0]).all(), "padding rows must not be written"


@pytest.mark.parametrize("R,S,Cc", [(5, 16, 192), (7, 8, 320), (150, 16, 64), (1, 8, 64), (150, 16, 192)])   # the last one runs with a pinned N tile per CTA
def test_conv_bf16_padded_to_flat_gathered_source(lib, R, S, Cc):
    """1x1 conv from a PADDED source to FLAT rows (qkv projection / data gradient of the attention output projection):
    TMA gathers the valid pixels, so NaN padding rows of the source must not leak; q|k panels leave through the staged
    epilogue, V transposed; the accumulate-into-existing-gradient port (residual, FLAT) is covered too."""
    from view_fusion_b200 import ops
    torch.manual_seed(12)
    dtype = torch.bfloat16
    rnd = lambda *s: bf16r(torch.randn(*s))
    x, w = rnd(R, Cc, S, S), bf16r(rnd(3 * Cc, Cc, 1, 1) / math.sqrt(Cc))
    ref = F.conv2d(x, w)
    wp = ops.pack_conv_weight(w.cuda(), dtype)
    src = ops.to_padded(x, dtype, fill=float("nan")).cuda()
    out, vt = ops.conv2d([src], [1], wp, R, S, S, 3 * Cc, qkv_split=Cc, out_padded=False)
    got = ops.from_nhwc(out, R, S, S).cpu()
    assert rel(got[:, : 2 * Cc], ref[:, : 2 * Cc]) < 4e-3
    assert rel(vt.float().cpu().view(R, Cc, S, S), ref[:, 2 * Cc:]) < 4e-3
    # plain FLAT output + FLAT residual
    res = rnd(R, 3 * Cc, S, S)
    out2 = ops.conv2d([src], [1], wp, R, S, S, 3 * Cc, out_padded=False, residual=ops.to_nhwc(res, dtype).cuda())
    assert rel(ops.from_nhwc(out2, R, S, S).cpu(), ref + res) < 4e-3


@pytest.mark.parametrize("R,S,Cc", [(3, 32, 64), (40, 16, 128), (150, 8, 192), (2, 4, 64)])
def test_conv_bf16_stride2_gathers_pixel_phases(lib, R, S, Cc):
    """Downsample (3x3, stride 2, unet.py:195-201) over the OUTPUT pixels only: per-tap TMA gathers of the input's pixel
    phases, zero halo from out-of-bounds fill (NaN padding rows of the source must not leak), fused GroupNorm sums.
    S is the OUTPUT size; (2, 4, 64) does not fit the gather (4x4 outputs) and checks the full-resolution fallback."""
    from view_fusion_b200 import ops
    torch.manual_seed(13)
    dtype = torch.bfloat16
    rnd = lambda *s: bf16r(torch.randn(*s))
    x, w, bias = rnd(R, Cc, 2 * S, 2 * S), bf16r(rnd(Cc, Cc, 3, 3) / math.sqrt(9 * Cc)), torch.randn(Cc)
    ref = F.conv2d(x, w, bias, stride=2, padding=1)
    wp = ops.pack_conv_weight(w.cuda(), dtype)
    gathers = S >= 8
    src = ops.to_padded(x, dtype, fill=float("nan") if gathers else 0.0).cuda()
    out, stats = ops.conv2d([src], [3], wp, R, 2 * S, 2 * S, Cc, stride=2, bias=bias.cuda(), want_stats=True)
    got = ops.from_padded(out, R, S, S).cpu()
    assert rel(got, ref) < 4e-3
    want = torch.stack([got.sum(dim=(2, 3)), (got * got).sum(dim=(2, 3))], dim=-1)
    assert rel(stats, want) < 1e-5


def test_conv_bf16_final_layer_fp32_out(lib):
    from view_fusion_b200 import ops
    R, S, segs, cout = 2, 16, [(64, 3)], 6
    ref, *t = _conv_case(torch.bfloat16, R, S, segs, cout, 1, False, False)
    out = _run_conv(torch.bfloat16, R, S, segs, cout, 1, *t, out_dtype=torch.float32, out_ld=8, cout_pad=16, out_padded=False)
    got = out.view(R, S, S, 8)[..., :6].permute(0, 3, 1, 2).cpu()
    assert rel(got, ref) < 1e-5


def test_conv_bf16_qkv_split_writes_v_transposed(lib):
    from view_fusion_b200 import ops
    R, S, Cc = 3, 8, 128
    ref, *t = _conv_case(torch.bfloat16, R, S, [(Cc, 1)], 3 * Cc, 1, False, False)
    out, vt = _run_conv(torch.bfloat16, R, S, [(Cc, 1)], 3 * Cc, 1, *t, qkv_split=Cc, out_padded=False)
    got = ops.from_nhwc(out, R, S, S)
    assert rel(got[:, : 2 * Cc], ref[:, : 2 * Cc]) < 4e-3
    assert rel(vt.float().cpu().view(R, Cc, S, S), ref[:, 2 * Cc:]) < 4e-3


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("R,L,Cc", [(3, 64, 128), (2, 256, 192), (1, 64, 320), (2, 256, 64)])
def test_attention(lib, dtype, R, L, Cc):
    from view_fusion_b200 import ops
    torch.manual_seed(3)
    qkv = torch.randn(R, L, 3 * Cc) * 1.3
    if dtype == torch.bfloat16:
        qkv = bf16r(qkv)
    q, k, v = qkv[..., :Cc], qkv[..., Cc:2 * Cc], qkv[..., 2 * Cc:]
    p = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(Cc), dim=-1)
    ref = (p @ v).reshape(R * L, Cc)
    qd = qkv.reshape(R * L, 3 * Cc).to(dtype).cuda()
    vt = v.transpose(1, 2).contiguous().to(dtype).cuda() if dtype == torch.bfloat16 else None
    out = ops.attention(qd, vt, R, L, Cc)
    torch.cuda.synchronize()
    assert rel(out, ref) < (1e-2 if dtype == torch.bfloat16 else 5e-6)
    if dtype == torch.bfloat16:
        ops.force_simt(True)
        try:
            simt = ops.attention(qd, vt, R, L, Cc)
        finally:
            ops.force_simt(False)
        assert rel(out, simt) < 1e-2
        # inference variant: no transposed copy, V is read row-major from qkv as an MN-major tensor-core operand
        out_mn = ops.attention(qd, None, R, L, Cc)
        torch.cuda.synchronize()
        assert rel(out_mn, ref) < 1e-2
        assert rel(out_mn, out) < 2e-3


def test_embed_table(lib):
    from view_fusion_b200 import ops
    cfg = O.SMALL_V100
    sd = O.init_state_dict(cfg, 4)
    rows = 5
    level = torch.rand(rows, 1) * 0.999 + 1e-4
    angle = (2 * math.pi / 24) * torch.randint(0, 24, (rows, 1)).float()
    t = O.time_embedding(sd, cfg, angle, level)
    names = [n for n, _, _ in O.param_shapes(cfg) if n.endswith("noise_func.noise_func.0.weight")]
    ew = torch.cat([sd[n] for n in names])
    eb = torch.cat([sd[n[:-6] + "bias"] for n in names])
    ref = F.linear(t, ew, eb).squeeze(1)
    got = ops.embed(level.reshape(-1).cuda(), angle.reshape(-1).cuda(), cfg["inner_channel"], sd["noise_level_mlp.0.weight"].cuda(),
                    sd["noise_level_mlp.0.bias"].cuda(), sd["noise_level_mlp.2.weight"].cuda(), sd["noise_level_mlp.2.bias"].cuda(),
                    ew.cuda(), eb.cuda())
    assert ew.shape[0] == 5568
    assert rel(got, ref) < 5e-6


def test_upsample_and_pack(lib):
    from view_fusion_b200 import _lib, ops
    x = torch.randn(2, 64, 8, 8)
    up = ops.upsample2x(ops.to_padded(x, fill=float("nan")).cuda(), 2, 8, 8)
    assert torch.equal(ops.from_padded(up, 2, 16, 16).cpu(), F.interpolate(x, scale_factor=2, mode="nearest"))
    assert float(up.view(2, 17, 17, 64)[:, 0].abs().max()) == 0.0 and float(up.view(2, 17, 17, 64)[:, :, 0].abs().max()) == 0.0
    flat = ops.to_nhwc(x).cuda()
    assert torch.equal(ops.padded_to_flat(ops.flat_to_padded(flat, 2, 8, 8), 2, 8, 8), flat)
    # pack_views == stack_views + im2col of the first conv
    bt = O.synthetic_batch(3, 4, 16, seed=9, ragged=True, nmax=5)
    y_t = torch.randn(3, 3, 16, 16)
    xs, _, _ = O.stack_views(bt["y_cond"], y_t, bt["view_count"], bt["angle"], torch.zeros(3, 1))
    R = xs.shape[0]
    cols = F.unfold(xs, 3, padding=1).view(R, 6, 9, 256).permute(0, 3, 2, 1).reshape(R * 256, 54)   # k = tap*6 + c
    off = torch.zeros(4, dtype=torch.int32)
    off[1:] = torch.cumsum(bt["view_count"], 0)
    x0 = torch.empty(R * 256, 64, device="cuda")
    img = torch.empty(R, dtype=torch.int32, device="cuda")
    yc, yt, offd = bt["y_cond"].cuda(), y_t.cuda(), off.cuda()
    _lib.check(lib.vf_pack_views(yc.data_ptr(), yt.data_ptr(), offd.data_ptr(), 3, 5, 3, 16, 16, R, 64, _lib.VF_F32, x0.data_ptr(),
                                 img.data_ptr(), _lib.stream_handle()), "pack")
    assert torch.equal(x0.cpu()[:, :54], cols) and float(x0[:, 54:].abs().sum()) == 0.0
    assert img.cpu().tolist() == sum(([b] * int(v) for b, v in enumerate(bt["view_count"])), [])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_conv_fused_groupnorm_statistics(lib, dtype):
    """vf_conv_args::stats == per (image, channel) sum / sum-of-squares of the stored output."""
    from view_fusion_b200 import ops
    R, S, segs, cout = 3, 16, [(64, 3)], 128
    ref, *t = _conv_case(dtype, R, S, segs, cout, 1, True, True)
    out, stats = _run_conv(dtype, R, S, segs, cout, 1, *t, want_stats=True)
    o = ops.from_padded(out, R, S, S).cpu()               # stored values (bf16-rounded in bf16 mode)
    want = torch.stack([o.sum(dim=(2, 3)), (o * o).sum(dim=(2, 3))], dim=-1)
    assert rel(stats, want) < 1e-5
    # and GroupNorm driven by the fused statistics == GroupNorm of the stored tensor
    gamma, beta = torch.rand(cout) + 0.5, torch.randn(cout) * 0.1
    y = ops.gn_apply(out, None, R, S, S, 32, stats, gamma.cuda(), beta.cuda(), True)
    refn = O.swish(F.group_norm(o, 32, gamma, beta, eps=1e-5))
    assert rel(ops.from_padded(y, R, S, S), refn) < (4e-3 if dtype == torch.bfloat16 else 2e-6)


def test_probe_shifted_umma_descriptor(lib):
    """Hardware probe: tcgen05 smem descriptor shifted by whole 128-byte rows inside a TMA-written SWIZZLE_128B tile."""
    from view_fusion_b200 import _lib
    torch.manual_seed(0)
    A = bf16r(torch.randn(512, 64))
    B = bf16r(torch.randn(64, 64))
    Ad, Bd = A.to(torch.bfloat16).cuda(), B.to(torch.bfloat16).cuda()
    res = {}
    for shift in (0, 8, 64, 1, 3, 7, 9, 65, 67):
        for bo in (0, 1):
            out = torch.zeros(128, 64, device="cuda")
            _lib.check(_lib.load_probes().vf_debug_umma_shift(Ad.data_ptr(), 512, Bd.data_ptr(), shift, bo, out.data_ptr(), _lib.stream_handle()), "probe")
            torch.cuda.synchronize()
            ref = A[shift:shift + 128] @ B.t()
            res[(shift, bo)] = rel(out, ref)
    print("\nUMMA_SHIFT_PROBE", {k: round(v, 5) for k, v in res.items()})
    assert res[(0, 0)] < 1e-5 and res[(8, 0)] < 1e-5 and res[(64, 0)] < 1e-5     # atom-aligned shifts must work

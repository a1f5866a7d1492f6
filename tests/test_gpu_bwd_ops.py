"""GPU: stage-level BACKWARD operators through the C ABI against fp64 autograd on the same (bf16-rounded) operands.

The whole-model gradient tests (tests/test_gpu_train.py) compare against the reference's autograd through ~60 layers,
where bf16 rounding dominates; these tests pin every backward kernel on its own so that a wrong kernel cannot hide inside
that tolerance: GroupNorm(+Swish) backward, the tcgen05 data gradients (transposed tap-flipped packs; stride 2 through zero
insertion; the two 1x1 row-order modes of the attention block), the up-sampling backward, the bias / embedding column sums
and the embedding-MLP backward.  Reference: autograd of model/unet.py:147-277 (experiment.py:292 loss.backward()).
"""
import ctypes as C
import math

import pytest
import torch
import torch.nn.functional as F

import vf_oracle as O
from gpu_util import bf16r, margin, rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from view_fusion_b200 import _lib
    return _lib.require_device()


def _dt(dtype):
    from view_fusion_b200 import _lib
    return _lib.VF_BF16 if dtype == torch.bfloat16 else _lib.VF_F32


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("C0,C1,S,swish,acc", [(64, 0, 16, True, False), (192, 128, 8, True, True), (128, 64, 32, True, False),
                                               (320, 320, 8, True, False), (192, 0, 16, False, True), (64, 0, 64, True, False)])
@pytest.mark.parametrize("one_pass", [True, False])
def test_groupnorm_swish_backward(lib, dtype, C0, C1, S, swish, acc, one_pass):
    """one_pass: the (image, slab of groups)-resident kernel where the slab fits in shared memory; False forces the two-pass
    reduce + apply kernels (vf_debug_flags 0x2000), which otherwise only run for the shapes that do not fit."""
    from view_fusion_b200 import _lib, ops
    torch.manual_seed(C0 + C1 + S)
    R, groups = 3, 32
    rnd = (lambda *s: bf16r(torch.randn(*s))) if dtype == torch.bfloat16 else (lambda *s: torch.randn(*s))
    x0 = rnd(R, C0, S, S) * 1.5 + 0.3
    x1 = rnd(R, C1, S, S) * 0.7 - 0.2 if C1 else None
    x0 = bf16r(x0) if dtype == torch.bfloat16 else x0
    x1 = bf16r(x1) if (dtype == torch.bfloat16 and x1 is not None) else x1
    Cc = C0 + C1
    gamma, beta = torch.rand(Cc) + 0.5, torch.randn(Cc) * 0.1
    dy = rnd(R, Cc, S, S)
    old0, old1 = rnd(R, C0, S, S), (rnd(R, C1, S, S) if C1 else None)
    # fp64 autograd reference
    a0 = x0.double().requires_grad_(True)
    a1 = x1.double().requires_grad_(True) if C1 else None
    gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    xc = a0 if a1 is None else torch.cat([a0, a1], 1)
    y = F.group_norm(xc, groups, gd, bd, eps=1e-5)
    y = y * torch.sigmoid(y) if swish else y
    y.backward(dy.double())
    # device
    s0 = ops.to_padded(x0, dtype, fill=float("nan")).cuda()
    s1 = ops.to_padded(x1, dtype, fill=float("nan")).cuda() if C1 else None
    st = ops.gn_stats(s0, s1, R, S, S)
    dyp = ops.to_padded(dy, dtype, fill=float("nan")).cuda()
    P = (S + 1) * (S + 1)
    dx0 = (ops.to_padded(old0, dtype) if acc else torch.full((R * P, C0), float("nan")).to(dtype)).cuda()
    dx1 = None
    if C1:
        dx1 = (ops.to_padded(old1, dtype) if acc else torch.full((R * P, C1), float("nan")).to(dtype)).cuda()
    scratch = torch.empty(R * Cc * 2, device="cuda")
    dgm, dbt = torch.full((Cc,), 0.5, device="cuda"), torch.full((Cc,), -0.25, device="cuda")       # accumulated into
    gm, bt = gamma.cuda(), beta.cuda()
    lib.vf_debug_flags(0 if one_pass else 0x2000)
    try:
        _lib.check(lib.vf_gn_backward(s0.data_ptr(), C0, st.data_ptr(), Cc, _lib.ptr(s1), C1, st.data_ptr() + 8 * C0 if C1 else 0, Cc, _dt(dtype),
                                      R, S, S, groups, gm.data_ptr(), bt.data_ptr(), int(swish), dyp.data_ptr(), scratch.data_ptr(), dgm.data_ptr(),
                                      dbt.data_ptr(), dx0.data_ptr(), int(acc), _lib.ptr(dx1), int(acc), None, _lib.stream_handle()), "vf_gn_backward")
        torch.cuda.synchronize()
    finally:
        lib.vf_debug_flags(0)
    tag = (f"gn_backward {'bf16' if dtype == torch.bfloat16 else 'fp32'} C={C0}+{C1} {S}x{S} swish={int(swish)} acc={int(acc)} "
           f"{'one-pass' if one_pass else 'two-pass'}")
    tol_x, tol_p = (1e-2, 5e-3) if dtype == torch.bfloat16 else (2e-5, 2e-5)
    ref0 = a0.grad.float() + (old0 if acc else 0)
    ok = margin(f"{tag}: dx0 rel-L2 vs fp64 autograd", rel(ops.from_padded(dx0, R, S, S), ref0), tol_x)
    if C1:
        ref1 = a1.grad.float() + (old1 if acc else 0)
        ok &= margin(f"{tag}: dx1 rel-L2 vs fp64 autograd", rel(ops.from_padded(dx1, R, S, S), ref1), tol_x)
    ok &= margin(f"{tag}: dgamma rel-L2", rel(dgm - 0.5, gd.grad.float()), tol_p)
    ok &= margin(f"{tag}: dbeta rel-L2", rel(dbt + 0.25, bd.grad.float()), tol_p)
    if not acc:      # gradients of padding rows are exact zeros (the weight-gradient GEMMs read them)
        pad = dx0.float().view(R, S + 1, S + 1, C0)
        assert float(pad[:, 0].abs().max()) == 0.0 and float(pad[:, :, 0].abs().max()) == 0.0
    assert ok


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("C0,S", [(64, 16), (320, 8), (128, 32)])
def test_groupnorm_deferred_bias_and_embedding(lib, dtype, C0, S):
    """vf_gn_shift: ResnetBlock.block1's bias and FeatureWiseAffine add (unet.py:243, :176) are never added in memory; the
    GroupNorm of block2 (forward and backward) acts on h + bias[c] + emb[img_row[img]][c] in closed form.  Checked against
    fp64 autograd of GroupNorm+Swish applied to the explicitly shifted tensor."""
    from view_fusion_b200 import _lib, ops
    torch.manual_seed(C0 + S)
    R, groups, rows, E, col = 4, 32, 3, C0 + 40, 8
    rnd = (lambda *s: bf16r(torch.randn(*s))) if dtype == torch.bfloat16 else (lambda *s: torch.randn(*s))
    h = rnd(R, C0, S, S) * 1.3 + 0.2
    h = bf16r(h) if dtype == torch.bfloat16 else h
    bias, emb = torch.randn(C0) * 0.5, torch.randn(rows, E)
    img_row = torch.randint(0, rows, (R,), dtype=torch.int32)
    gamma, beta = torch.rand(C0) + 0.5, torch.randn(C0) * 0.1
    dy = rnd(R, C0, S, S)
    shift_ref = (bias[None, :] + emb[img_row.long(), col:col + C0]).double()
    a0 = h.double().requires_grad_(True)
    gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    y = F.group_norm(a0 + shift_ref[:, :, None, None], groups, gd, bd, eps=1e-5)
    y = y * torch.sigmoid(y)
    y.backward(dy.double())
    s0 = ops.to_padded(h, dtype, fill=float("nan")).cuda()
    st = ops.gn_stats(s0, None, R, S, S)                      # raw sums of the STORED tensor (what the conv epilogue emits)
    bd_, ed_, ir_ = bias.cuda(), emb.cuda(), img_row.cuda()
    sh = ops.gn_shift(bd_, ed_[:, col:], ir_)
    sh.emb_ld = E
    gm, bt = gamma.cuda(), beta.cuda()
    out = ops.gn_apply(s0, None, R, S, S, groups, st, gm, bt, True, shift=sh)
    tag = f"gn deferred bias+emb {'bf16' if dtype == torch.bfloat16 else 'fp32'} C={C0} {S}x{S}"
    ok = margin(f"{tag}: forward rel-L2 vs fp64", rel(ops.from_padded(out, R, S, S), y.detach().float()), 4e-3 if dtype == torch.bfloat16 else 5e-6)
    dyp = ops.to_padded(dy, dtype, fill=float("nan")).cuda()
    dx0 = torch.full((R * (S + 1) * (S + 1), C0), float("nan")).to(dtype).cuda()
    scratch = torch.empty(R * C0 * 2, device="cuda")
    dgm, dbt = torch.zeros(C0, device="cuda"), torch.zeros(C0, device="cuda")
    _lib.check(lib.vf_gn_backward(s0.data_ptr(), C0, st.data_ptr(), C0, 0, 0, 0, 0, _dt(dtype), R, S, S, groups, gm.data_ptr(), bt.data_ptr(), 1,
                                  dyp.data_ptr(), scratch.data_ptr(), dgm.data_ptr(), dbt.data_ptr(), dx0.data_ptr(), 0, 0, 0, C.byref(sh),
                                  _lib.stream_handle()), "vf_gn_backward")
    torch.cuda.synchronize()
    tol_x, tol_p = (1e-2, 5e-3) if dtype == torch.bfloat16 else (2e-5, 2e-5)
    ok &= margin(f"{tag}: dx rel-L2 vs fp64 autograd", rel(ops.from_padded(dx0, R, S, S), a0.grad.float()), tol_x)
    ok &= margin(f"{tag}: dgamma rel-L2", rel(dgm, gd.grad.float()), tol_p)
    ok &= margin(f"{tag}: dbeta rel-L2", rel(dbt, bd.grad.float()), tol_p)
    assert ok


# ---------------------------------------------------------------------------------------------------------------------
def _pack_t(lib, w, dtype, n_stride=None):
    """Transposed, tap-flipped pack of an OIHW weight: [cin, taps * n_stride]."""
    from view_fusion_b200 import _lib
    cout, cin, k, _ = w.shape
    n_stride = cout if n_stride is None else n_stride
    dst = torch.zeros(cin, k * k * n_stride, dtype=dtype, device="cuda")
    wd = w.contiguous().cuda()
    _lib.check(lib.vf_pack_conv_weight_t(wd.data_ptr(), cout, cin, k, _dt(dtype), dst.data_ptr(), cin, k * k * n_stride, 0, n_stride,
                                         _lib.stream_handle()), "vf_pack_conv_weight_t")
    return dst


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("R,S,cin,cout,acc", [(3, 16, 64, 128, False), (2, 32, 128, 64, True), (5, 8, 320, 320, False), (2, 64, 64, 64, True),
                                              (40, 16, 192, 192, False)])
def test_dgrad_3x3_transposed_pack(lib, dtype, R, S, cin, cout, acc):
    """dX of a 3x3 stride-1 convolution = forward vf_conv2d over dY with the transposed, tap-flipped pack (+ accumulate port)."""
    from view_fusion_b200 import ops
    torch.manual_seed(R * S + cin)
    rnd = (lambda *s: bf16r(torch.randn(*s))) if dtype == torch.bfloat16 else (lambda *s: torch.randn(*s))
    w = rnd(cout, cin, 3, 3) / math.sqrt(9 * cin)
    w = bf16r(w) if dtype == torch.bfloat16 else w
    dy, old = rnd(R, cout, S, S), rnd(R, cin, S, S)
    x = torch.zeros(R, cin, S, S, dtype=torch.float64, requires_grad=True)
    F.conv2d(x, w.double(), padding=1).backward(dy.double())
    ref = x.grad.float() + (old if acc else 0)
    wt = _pack_t(lib, w, dtype)
    out = ops.conv2d([ops.to_padded(dy, dtype).cuda()], [3], wt, R, S, S, cin,
                     residual=ops.to_padded(old, dtype, fill=float("nan")).cuda() if acc else None)
    torch.cuda.synchronize()
    tag = f"dgrad 3x3 {'bf16 tcgen05' if dtype == torch.bfloat16 else 'fp32'} R={R} {S}x{S} {cout}->{cin} acc={int(acc)}"
    assert margin(f"{tag}: dX rel-L2 vs fp64 autograd", rel(ops.from_padded(out, R, S, S), ref), 4e-3 if dtype == torch.bfloat16 else 5e-6)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("R,S,Cc", [(3, 16, 64), (2, 8, 128), (40, 8, 192)])
def test_dgrad_stride2_zero_insertion(lib, dtype, R, S, Cc):
    """Downsample (3x3, stride 2) data gradient: dY (S x S) scattered onto the even pixels of a zero 2S x 2S grid, then the
    stride-1 transposed convolution at source resolution (what conv_backward runs for unet.py:195-201)."""
    from view_fusion_b200 import _lib, ops
    torch.manual_seed(S + Cc)
    rnd = (lambda *s: bf16r(torch.randn(*s))) if dtype == torch.bfloat16 else (lambda *s: torch.randn(*s))
    w = rnd(Cc, Cc, 3, 3) / math.sqrt(9 * Cc)
    w = bf16r(w) if dtype == torch.bfloat16 else w
    dy = rnd(R, Cc, S, S)
    x = torch.zeros(R, Cc, 2 * S, 2 * S, dtype=torch.float64, requires_grad=True)
    F.conv2d(x, w.double(), stride=2, padding=1).backward(dy.double())
    dyp = ops.to_padded(dy, dtype, fill=float("nan")).cuda()
    z = torch.full((R * (2 * S + 1) * (2 * S + 1), Cc), float("nan"), dtype=dtype, device="cuda")
    _lib.check(lib.vf_zero_insert2x(dyp.data_ptr(), _dt(dtype), R, S, S, Cc, z.data_ptr(), _lib.stream_handle()), "vf_zero_insert2x")
    out = ops.conv2d([z], [3], _pack_t(lib, w, dtype), R, 2 * S, 2 * S, Cc)
    torch.cuda.synchronize()
    tag = f"dgrad 3x3 stride 2 {'bf16 tcgen05' if dtype == torch.bfloat16 else 'fp32'} R={R} out {S}x{S} C={Cc}"
    assert margin(f"{tag}: dX rel-L2 vs fp64 autograd", rel(ops.from_padded(out, R, 2 * S, 2 * S), x.grad.float()),
                  4e-3 if dtype == torch.bfloat16 else 5e-6)


@pytest.mark.parametrize("R,S,Cc", [(5, 16, 192), (7, 8, 320), (150, 16, 64)])
def test_dgrad_1x1_row_order_modes_bf16(lib, R, S, Cc):
    """The attention block's projections change the row order, so do their data gradients:
       qkv (PADDED -> FLAT forward): dN (PADDED) = dQKV (FLAT) W       -> FLAT source, line-map epilogue, accumulate port
       out (FLAT -> PADDED forward): dO (FLAT)   = dOut (PADDED) W     -> gathered PADDED source (NaN padding must not leak)"""
    from view_fusion_b200 import ops
    torch.manual_seed(S * Cc + R)
    dtype = torch.bfloat16
    rnd = lambda *s: bf16r(torch.randn(*s))
    # qkv
    w = bf16r(rnd(3 * Cc, Cc, 1, 1) / math.sqrt(Cc))
    dqkv, old = rnd(R, 3 * Cc, S, S), rnd(R, Cc, S, S)
    ref = torch.einsum("rnhw,nc->rchw", dqkv.double(), w.double()[:, :, 0, 0]).float() + old
    out = ops.conv2d([ops.to_nhwc(dqkv, dtype).cuda()], [1], _pack_t(lib, w, dtype), R, S, S, Cc, in_padded=False, out_padded=True,
                     residual=ops.to_padded(old, dtype, fill=float("nan")).cuda())
    ok = margin(f"dgrad 1x1 qkv bf16 tcgen05 R={R} {S}x{S} C={Cc}: dN rel-L2 vs fp64", rel(ops.from_padded(out, R, S, S), ref), 4e-3)
    # attention out-projection
    w2 = bf16r(rnd(Cc, Cc, 1, 1) / math.sqrt(Cc))
    dout = rnd(R, Cc, S, S)
    ref2 = torch.einsum("rnhw,nc->rchw", dout.double(), w2.double()[:, :, 0, 0]).float()
    out2 = ops.conv2d([ops.to_padded(dout, dtype, fill=float("nan")).cuda()], [1], _pack_t(lib, w2, dtype), R, S, S, Cc, in_padded=True,
                      out_padded=False)
    torch.cuda.synchronize()
    ok &= margin(f"dgrad 1x1 attn.out bf16 tcgen05 R={R} {S}x{S} C={Cc}: dO rel-L2 vs fp64", rel(ops.from_nhwc(out2, R, S, S), ref2), 4e-3)
    assert ok


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("acc", [False, True])
def test_upsample2x_backward(lib, dtype, acc):
    from view_fusion_b200 import _lib, ops
    torch.manual_seed(3)
    R, Cc, S = 3, 128, 8
    rnd = (lambda *s: bf16r(torch.randn(*s))) if dtype == torch.bfloat16 else (lambda *s: torch.randn(*s))
    dy, old = rnd(R, Cc, 2 * S, 2 * S), rnd(R, Cc, S, S)
    ref = 4.0 * F.avg_pool2d(dy.double(), 2).float() + (old if acc else 0)
    dyp = ops.to_padded(dy, dtype, fill=float("nan")).cuda()
    dx = (ops.to_padded(old, dtype) if acc else torch.full((R * (S + 1) * (S + 1), Cc), float("nan")).to(dtype)).cuda()
    _lib.check(lib.vf_upsample2x_backward(dyp.data_ptr(), _dt(dtype), R, S, S, Cc, dx.data_ptr(), int(acc), _lib.stream_handle()), "upsample_bwd")
    torch.cuda.synchronize()
    tag = f"upsample2x_backward {'bf16' if dtype == torch.bfloat16 else 'fp32'} acc={int(acc)}"
    assert margin(f"{tag}: dx rel-L2 vs fp64", rel(ops.from_padded(dx, R, S, S), ref), 4e-3 if dtype == torch.bfloat16 else 1e-6)
    if not acc:
        pad = dx.float().view(R, S + 1, S + 1, Cc)
        assert float(pad[:, 0].abs().max()) == 0.0 and float(pad[:, :, 0].abs().max()) == 0.0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("R,S,cout,ld", [(5, 16, 192, 192), (3, 64, 64, 64), (7, 8, 320, 320), (2, 32, 6, 64)])
def test_colsum_bias_and_embedding_gradient(lib, dtype, R, S, cout, ld):
    """db (+ a second bias) and the embedding-table gradient = per-image column sums of dY (unet.py:176, :214 backward)."""
    from view_fusion_b200 import _lib, ops
    torch.manual_seed(S + cout)
    rnd = (lambda *s: bf16r(torch.randn(*s))) if dtype == torch.bfloat16 else (lambda *s: torch.randn(*s))
    dy = torch.zeros(R, ld, S, S)
    dy[:, :cout] = rnd(R, cout, S, S)
    dyp = ops.to_padded(dy, dtype).cuda()                     # padding rows hold zeros (contract)
    rows, E, col = 3, cout + 24, 16
    img_row = torch.randint(0, rows, (R,), dtype=torch.int32)
    db0, db1 = torch.full((cout,), 1.0, device="cuda"), torch.full((cout,), -2.0, device="cuda")
    demb = torch.zeros(rows, E, device="cuda")
    ir = img_row.cuda()
    _lib.check(lib.vf_colsum_bias(dyp.data_ptr(), _dt(dtype), R, (S + 1) * (S + 1), ld, cout, db0.data_ptr(), db1.data_ptr(), demb.data_ptr(),
                                  ir.data_ptr(), E, col, _lib.stream_handle()), "vf_colsum_bias")
    torch.cuda.synchronize()
    per_img = dy[:, :cout].double().sum(dim=(2, 3))
    ref_b = per_img.sum(0).float()
    ref_e = torch.zeros(rows, E, dtype=torch.float64)
    ref_e.index_add_(0, img_row.long(), F.pad(per_img, (col, E - col - cout)))
    tag = f"colsum_bias {'bf16' if dtype == torch.bfloat16 else 'fp32'} R={R} {S}x{S} cout={cout} ld={ld}"
    ok = margin(f"{tag}: db rel-L2 vs fp64", rel(db0 - 1.0, ref_b), 2e-5)
    ok &= margin(f"{tag}: second bias rel-L2", rel(db1 + 2.0, ref_b), 2e-5)
    ok &= margin(f"{tag}: demb rel-L2", rel(demb, ref_e.float()), 2e-5)
    assert float(demb[:, :col].abs().max()) == 0.0 and float(demb[:, col + cout:].abs().max()) == 0.0
    assert ok


def test_embed_backward(lib):
    """Backward of PositionalEncoding -> noise_level_mlp -> the 30 per-block Linears (unet.py:115-116, 27-32, 165-176)."""
    from view_fusion_b200 import _lib
    cfg = O.SMALL_V100
    ic = cfg["inner_channel"]
    sd = {k: v.double() for k, v in O.init_state_dict(cfg, 4).items()}
    rows = 5
    torch.manual_seed(8)
    level = (torch.rand(rows, 1) * 0.999 + 1e-4)
    angle = (2 * math.pi / 24) * torch.randint(0, 24, (rows, 1)).float()
    names = [n for n, _, _ in O.param_shapes(cfg) if n.endswith("noise_func.noise_func.0.weight")]
    ew = torch.cat([sd[n] for n in names]).requires_grad_(True)
    eb = torch.cat([sd[n[:-6] + "bias"] for n in names]).requires_grad_(True)
    mlp = {k: sd[k].clone().requires_grad_(True) for k in ("noise_level_mlp.0.weight", "noise_level_mlp.0.bias", "noise_level_mlp.2.weight",
                                                           "noise_level_mlp.2.bias")}
    t = O.time_embedding(mlp, cfg, angle.double(), level.double())
    emb = F.linear(t, ew, eb).squeeze(1)
    E = emb.shape[1]
    demb = torch.randn(rows, E)
    emb.backward(demb.double())
    f = lambda x: x.detach().float().contiguous().cuda()
    w0, b0, w2, b2 = (f(mlp[k]) for k in ("noise_level_mlp.0.weight", "noise_level_mlp.0.bias", "noise_level_mlp.2.weight", "noise_level_mlp.2.bias"))
    ewd, dembd = f(ew), demb.cuda()
    lv, an = level.reshape(-1).cuda(), angle.reshape(-1).cuda()
    rowbuf = torch.empty(rows * 11 * ic, device="cuda")
    dew, deb = torch.full((E, ic), float("nan"), device="cuda"), torch.full((E,), float("nan"), device="cuda")
    dw0, db0, dw2, db2 = (torch.zeros_like(x) for x in (w0, b0, w2, b2))
    _lib.check(lib.vf_embed_backward(lv.data_ptr(), an.data_ptr(), rows, ic, w0.data_ptr(), b0.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                                     ewd.data_ptr(), E, dembd.data_ptr(), rowbuf.data_ptr(), dew.data_ptr(), deb.data_ptr(), dw0.data_ptr(),
                                     db0.data_ptr(), dw2.data_ptr(), db2.data_ptr(), _lib.stream_handle()), "vf_embed_backward")
    torch.cuda.synchronize()
    ok = True
    for name, got, ref in (("dEw", dew, ew.grad), ("dEb", deb, eb.grad), ("dW0", dw0, mlp["noise_level_mlp.0.weight"].grad),
                           ("db0", db0, mlp["noise_level_mlp.0.bias"].grad), ("dW2", dw2, mlp["noise_level_mlp.2.weight"].grad),
                           ("db2", db2, mlp["noise_level_mlp.2.bias"].grad)):
        ok &= margin(f"embed_backward fp32 rows={rows} E={E}: {name} rel-L2 vs fp64 autograd", rel(got, ref.float()), 1e-4)
    assert ok

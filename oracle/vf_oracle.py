"""CPU oracle for the ViewFusion hot path.

TEST INFRASTRUCTURE ONLY.  This module is a functional, fp32, CPU PyTorch
restatement of the reference algorithm (bronemos/view-fusion, `model/unet.py`
and `model/view_fusion.py`).  It exists so that the CUDA path can be checked
against something that runs everywhere.  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import it; the
product package (`view_fusion_b200/`) never does.

Parity pinning: the reference ships no tests, golden vectors or known-answer
values for this path, so the oracle is pinned against the
reference ITSELF: `oracle/make_golden.py` imports `/root/reference/model`
unmodified in the build container, loads the same deterministic state_dict into
it, and (a) asserts this restatement agrees with it, (b) writes the reference's
outputs to `tests/golden/*.npz`.  `tests/test_oracle_golden.py` re-checks the
oracle against those committed reference outputs on any machine.

Two helpers outside `model/` are restated as well: `process_batch_u8` (data/nmr_dataset.py:10-52,
checked against the literal per-sample steps in tests/test_oracle_golden.py) and `ssim` — the reference calls
pytorch-msssim 1.0.0, which is absent from this image, so that ONE function restates the package's published
algorithm and its parity is UNPINNED (cross-checked against a float64 brute-force evaluation of the definition only).

Everything is written functionally over a flat ``state_dict`` (reference key
layout) so it shares no module code with the product.
Each function cites the reference lines it restates.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

# --------------------------------------------------------------------------------------
# configuration / layout
# --------------------------------------------------------------------------------------

SMALL_V100 = dict(  # configs/small-v100.yaml:19-30
    in_channel=6, out_channel=6, inner_channel=64, norm_groups=32,
    channel_mults=(1, 2, 3, 5), attn_res=(16,), res_blocks=3, image_size=64,
)
TINY = dict(  # small enough for second-scale CPU tests; same topology features
    in_channel=6, out_channel=6, inner_channel=32, norm_groups=32,
    channel_mults=(1, 2, 3), attn_res=(8,), res_blocks=1, image_size=16,
)
BETA_TRAIN = dict(schedule="linear", num_timesteps=2000, linear_start=1e-6, linear_end=1e-2)


def unet_layout(cfg) -> dict:
    """Module table of the reference UNet (unet.py:38-112).

    Returns dict(downs=[...], mid=[...], ups=[...], final=(cin, cout)) where each
    entry is ("conv", cin, cout) | ("rb", cin, cout, attn) | ("down", c) | ("up", c).
    """
    ic = cfg["inner_channel"]
    mults = list(cfg["channel_mults"])
    attn_res = set(cfg["attn_res"])
    rb = cfg["res_blocks"]
    res = cfg["image_size"]
    pre = ic
    feat = [pre]
    downs: List[tuple] = [("conv", cfg["in_channel"], ic)]
    for i, m in enumerate(mults):
        last = i == len(mults) - 1
        for _ in range(rb):
            downs.append(("rb", pre, ic * m, res in attn_res))
            feat.append(ic * m)
            pre = ic * m
        if not last:
            downs.append(("down", pre))
            feat.append(pre)
            res //= 2
    mid = [("rb", pre, pre, True), ("rb", pre, pre, False)]
    ups: List[tuple] = []
    for i in reversed(range(len(mults))):
        last = i < 1
        for _ in range(rb + 1):
            ups.append(("rb", pre + feat.pop(), ic * mults[i], res in attn_res))
            pre = ic * mults[i]
        if not last:
            ups.append(("up", pre))
            res *= 2
    out_c = cfg["out_channel"] if cfg.get("out_channel") is not None else cfg["in_channel"]
    return dict(downs=downs, mid=mid, ups=ups, final=(pre, out_c))


def param_shapes(cfg) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(name, shape, kind) for every UNet tensor in reference state_dict order.

    kind in {"w", "b", "gn_w", "gn_b"}; fan-in for init is prod(shape[1:]).
    Order follows nn.Module registration order in unet.py (Appendix C).
    """
    ic = cfg["inner_channel"]
    G = cfg.get("norm_groups", 32)
    out: List[Tuple[str, Tuple[int, ...], str]] = []

    def lin(name, cin, cout):
        out.append((name + ".weight", (cout, cin), "w"))
        out.append((name + ".bias", (cout,), "b"))

    def conv(name, cin, cout, k, bias=True):
        out.append((name + ".weight", (cout, cin, k, k), "w"))
        if bias:
            out.append((name + ".bias", (cout,), "b"))

    def gn(name, c):
        out.append((name + ".weight", (c,), "gn_w"))
        out.append((name + ".bias", (c,), "gn_b"))

    def rb(name, cin, cout, attn):
        # registration order in ResnetBlock.__init__: noise_func, block1, block2, res_conv (unet.py:232-238)
        lin(name + ".res_block.noise_func.noise_func.0", ic, cout)
        gn(name + ".res_block.block1.block.0", cin)
        conv(name + ".res_block.block1.block.3", cin, cout, 3)
        gn(name + ".res_block.block2.block.0", cout)
        conv(name + ".res_block.block2.block.3", cout, cout, 3)
        if cin != cout:
            conv(name + ".res_block.res_conv", cin, cout, 1)
        if attn:
            gn(name + ".attn.norm", cout)
            conv(name + ".attn.qkv", cout, 3 * cout, 1, bias=False)
            conv(name + ".attn.out", cout, cout, 1)

    lin("noise_level_mlp.0", ic, 4 * ic)
    lin("noise_level_mlp.2", 4 * ic, ic)
    lay = unet_layout(cfg)
    for sec in ("downs", "mid", "ups"):
        for i, e in enumerate(lay[sec]):
            name = f"{sec}.{i}"
            if e[0] == "conv":
                conv(name, e[1], e[2], 3)
            elif e[0] == "rb":
                rb(name, e[1], e[2], e[3])
            else:  # down / up both hold `.conv`
                conv(name + ".conv", e[1], e[1], 3)
    cf, co = lay["final"]
    gn("final_conv.block.0", cf)
    conv("final_conv.block.3", cf, co, 3)
    return out


def init_state_dict(cfg, seed: int = 0, prefix: str = "", affine_jitter: float = 0.25) -> SD:
    """Deterministic, platform-independent random init (numpy PCG64, no transcendental).

    Mirrors torch's default Conv2d/Linear init law (kaiming-uniform a=sqrt(5) ==
    U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias).  GroupNorm affine gets
    a jitter around (1, 0) so that gamma/beta handling is actually exercised.
    """
    rng = np.random.default_rng(seed)
    sd: SD = {}
    for name, shape, kind in param_shapes(cfg):
        if kind == "w":
            bound = 1.0 / math.sqrt(int(np.prod(shape[1:])))
            a = rng.uniform(-bound, bound, size=shape)
            last_fan = int(np.prod(shape[1:]))
        elif kind == "b":
            bound = 1.0 / math.sqrt(last_fan)
            a = rng.uniform(-bound, bound, size=shape)
        elif kind == "gn_w":
            a = 1.0 + affine_jitter * rng.uniform(-1, 1, size=shape)
        else:
            a = affine_jitter * rng.uniform(-1, 1, size=shape)
        sd[prefix + name] = torch.from_numpy(a.astype(np.float32))
    return sd


# --------------------------------------------------------------------------------------
# schedule  (view_fusion.py:35-68, :321-362)
# --------------------------------------------------------------------------------------

def make_beta_schedule(schedule, num_timesteps, linear_start=1e-6, linear_end=1e-2, cosine_s=8e-3):
    """view_fusion.py:330-362 (float64 numpy)."""
    if schedule == "quad":
        return np.linspace(linear_start ** 0.5, linear_end ** 0.5, num_timesteps, dtype=np.float64) ** 2
    if schedule == "linear":
        return np.linspace(linear_start, linear_end, num_timesteps, dtype=np.float64)
    if schedule in ("warmup10", "warmup50"):
        frac = 0.1 if schedule == "warmup10" else 0.5
        b = linear_end * np.ones(num_timesteps, dtype=np.float64)
        n = int(num_timesteps * frac)
        b[:n] = np.linspace(linear_start, linear_end, n, dtype=np.float64)
        return b
    if schedule == "const":
        return linear_end * np.ones(num_timesteps, dtype=np.float64)
    if schedule == "jsd":
        return 1.0 / np.linspace(num_timesteps, 1, num_timesteps, dtype=np.float64)
    if schedule == "cosine":
        ts = np.arange(num_timesteps + 1, dtype=np.float64) / num_timesteps + cosine_s
        al = np.cos(ts / (1 + cosine_s) * math.pi / 2) ** 2
        al = al / al[0]
        return np.minimum(1 - al[1:] / al[:-1], 0.999)
    raise NotImplementedError(schedule)


def make_schedule(**beta_kw) -> SD:
    """The six fp32 buffers of set_new_noise_schedule (view_fusion.py:41-68)."""
    betas = make_beta_schedule(**beta_kw)
    alphas = 1.0 - betas
    gammas = np.cumprod(alphas, axis=0)
    gprev = np.append(1.0, gammas[:-1])
    var = betas * (1.0 - gprev) / (1.0 - gammas)
    f = lambda a: torch.tensor(a, dtype=torch.float32)
    return dict(
        gammas=f(gammas),
        sqrt_recip_gammas=f(np.sqrt(1.0 / gammas)),
        sqrt_recipm1_gammas=f(np.sqrt(1.0 / gammas - 1)),
        posterior_log_variance_clipped=f(np.log(np.maximum(var, 1e-20))),
        posterior_mean_coef1=f(betas * np.sqrt(gprev) / (1.0 - gammas)),
        posterior_mean_coef2=f((1.0 - gprev) * np.sqrt(alphas) / (1.0 - gammas)),
    )


# --------------------------------------------------------------------------------------
# UNet forward  (unet.py:114-138 and the blocks it calls)
# --------------------------------------------------------------------------------------

def positional_encoding(level: Tensor, dim: int) -> Tensor:
    """unet.py:147-157.  level (R,1) -> (R,1,dim): [sin(l*f_k), cos(l*f_k)], f_k = 1e4^(-k/count)."""
    count = dim // 2
    step = torch.arange(count, dtype=level.dtype, device=level.device) / count
    enc = level.unsqueeze(1) * torch.exp(-math.log(1e4) * step.unsqueeze(0))
    return torch.cat([torch.sin(enc), torch.cos(enc)], dim=-1)


def swish(x: Tensor) -> Tensor:
    return x * torch.sigmoid(x)  # unet.py:180-182


def time_embedding(sd: SD, cfg, angle: Tensor, time: Tensor, p: str = "") -> Tensor:
    """unet.py:115-116, :27-32.  -> (R,1,inner_channel)."""
    ic = cfg["inner_channel"]
    ta = torch.cat((positional_encoding(time, ic // 2), positional_encoding(angle, ic // 2)), dim=-1)
    h = F.linear(ta, sd[p + "noise_level_mlp.0.weight"], sd[p + "noise_level_mlp.0.bias"])
    return F.linear(swish(h), sd[p + "noise_level_mlp.2.weight"], sd[p + "noise_level_mlp.2.bias"])


def _block(sd: SD, name: str, x: Tensor, groups: int) -> Tensor:
    """Block.forward unet.py:207-218: GroupNorm -> Swish -> Conv3x3(pad 1)."""
    h = F.group_norm(x, groups, sd[name + ".block.0.weight"], sd[name + ".block.0.bias"], eps=1e-5)
    return F.conv2d(swish(h), sd[name + ".block.3.weight"], sd[name + ".block.3.bias"], padding=1)


def _attention(sd: SD, name: str, x: Tensor, groups: int) -> Tensor:
    """SelfAttention.forward unet.py:258-277 (single head, scale 1/sqrt(C), residual on pre-norm input)."""
    b, c, hh, ww = x.shape
    n = F.group_norm(x, groups, sd[name + ".norm.weight"], sd[name + ".norm.bias"], eps=1e-5)
    qkv = F.conv2d(n, sd[name + ".qkv.weight"])
    q, k, v = qkv.reshape(b, 3, c, hh * ww).unbind(1)            # (b, c, L) each
    s = torch.einsum("bcq,bck->bqk", q, k) / math.sqrt(c)
    pr = torch.softmax(s, dim=-1)
    o = torch.einsum("bqk,bck->bcq", pr, v).reshape(b, c, hh, ww)
    o = F.conv2d(o, sd[name + ".out.weight"], sd[name + ".out.bias"])
    return o + x


def _resblock(sd: SD, name: str, x: Tensor, t: Tensor, groups: int, attn: bool) -> Tensor:
    """ResnetBlocWithAttn.forward unet.py:299-303 -> ResnetBlock.forward :240-245."""
    rb = name + ".res_block"
    h = _block(sd, rb + ".block1", x, groups)
    e = F.linear(t, sd[rb + ".noise_func.noise_func.0.weight"], sd[rb + ".noise_func.noise_func.0.bias"])
    h = h + e.view(x.shape[0], -1, 1, 1)                          # FeatureWiseAffine additive branch :176
    h = _block(sd, rb + ".block2", h, groups)
    if (rb + ".res_conv.weight") in sd:
        x = F.conv2d(x, sd[rb + ".res_conv.weight"], sd[rb + ".res_conv.bias"])
    h = h + x
    return _attention(sd, name + ".attn", h, groups) if attn else h


def unet_forward(sd: SD, cfg, x: Tensor, angle: Tensor, time: Tensor, p: str = "",
                 taps: Optional[dict] = None) -> Tensor:
    """UNet.forward unet.py:114-138.  x (R,Cin,H,W), angle (R,1), time (R,1) -> (R,Cout,H,W).

    `taps`, if given, receives intermediate activations keyed by module name (for
    per-layer parity debugging of the CUDA path).
    """
    G = cfg.get("norm_groups", 32)
    lay = unet_layout(cfg)
    sub = {k[len(p):]: v for k, v in sd.items() if k.startswith(p)} if p else sd
    t = time_embedding(sub, cfg, angle, time)
    feats = []
    for i, e in enumerate(lay["downs"]):
        n = f"downs.{i}"
        if e[0] == "conv":
            x = F.conv2d(x, sub[n + ".weight"], sub[n + ".bias"], padding=1)
        elif e[0] == "rb":
            x = _resblock(sub, n, x, t, G, e[3])
        else:
            x = F.conv2d(x, sub[n + ".conv.weight"], sub[n + ".conv.bias"], stride=2, padding=1)  # :195-201
        feats.append(x)
        if taps is not None:
            taps[n] = x
    for i, e in enumerate(lay["mid"]):
        x = _resblock(sub, f"mid.{i}", x, t, G, e[3])
        if taps is not None:
            taps[f"mid.{i}"] = x
    for i, e in enumerate(lay["ups"]):
        n = f"ups.{i}"
        if e[0] == "rb":
            x = _resblock(sub, n, torch.cat((x, feats.pop()), dim=1), t, G, e[3])
        else:
            x = F.interpolate(x, scale_factor=2, mode="nearest")                           # :185-192
            x = F.conv2d(x, sub[n + ".conv.weight"], sub[n + ".conv.bias"], padding=1)
        if taps is not None:
            taps[n] = x
    return _block(sub, "final_conv", x, G)


# --------------------------------------------------------------------------------------
# view stacking, composition, DDPM step  (view_fusion.py:86-177, 216-300)
# --------------------------------------------------------------------------------------

def stack_views(y_cond: Tensor, y_t: Tensor, view_count: Tensor, angle: Tensor, level: Tensor):
    """view_fusion.py:95-115 / :244-263: first V_b views of each sample, target/level/angle repeated."""
    vc = [int(v) for v in view_count.tolist()]
    cond = torch.cat([y_cond[b, :v] for b, v in enumerate(vc)], dim=0)
    rep = torch.tensor(vc, dtype=torch.long, device=y_t.device)
    x = torch.cat([cond, torch.repeat_interleave(y_t, rep, dim=0)], dim=1)
    return x, torch.repeat_interleave(angle, rep, dim=0), torch.repeat_interleave(level, rep, dim=0)


def compose(out: Tensor, view_count: Tensor, weighting: bool = True):
    """view_fusion.py:116-150 / :265-296.

    out (sumV, 6|3, H, W) -> eps_hat (B,3,H,W), logits (sumV,3,H,W) un-padded | None,
    weights (B,maxV,3,H,W) zero in padded slots | None.  Softmax is over the view axis,
    per pixel and per channel.
    """
    vc = [int(v) for v in view_count.tolist()]
    B, mv = len(vc), max(vc)
    eps_all = out[:, :3]
    if not weighting:
        off = np.concatenate([[0], np.cumsum(vc)])
        eps = torch.stack([eps_all[off[b]:off[b + 1]].mean(dim=0) for b in range(B)])
        return eps, None, None
    logits = out[:, 3:]
    H, W = out.shape[-2:]
    lp = out.new_full((B, mv, 3, H, W), float("-inf"))
    ep = out.new_zeros((B, mv, 3, H, W))
    o = 0
    for b, v in enumerate(vc):
        lp[b, :v] = logits[o:o + v]
        ep[b, :v] = eps_all[o:o + v]
        o += v
    w = torch.softmax(lp, dim=1)
    return (ep * w).sum(dim=1), logits, w


def ddpm_update(sched: SD, y_t: Tensor, eps: Tensor, t: Tensor, z: Optional[Tensor], clip: bool = True) -> Tensor:
    """predict_start_from_noise :70-74, clamp :154-155, q_posterior :76-84, p_sample :176-177.

    z is the injected N(0,1) draw; the reference uses zeros when no t > 0 (:176).
    """
    g = lambda a: a.gather(-1, t).view(-1, 1, 1, 1)
    y0 = g(sched["sqrt_recip_gammas"]) * y_t - g(sched["sqrt_recipm1_gammas"]) * eps
    if clip:
        y0 = y0.clamp(-1.0, 1.0)
    mean = g(sched["posterior_mean_coef1"]) * y0 + g(sched["posterior_mean_coef2"]) * y_t
    lv = g(sched["posterior_log_variance_clipped"])
    if z is None or not bool((t > 0).any()):
        z = torch.zeros_like(y_t)
    return mean + z * (0.5 * lv).exp()


def p_sample(sd: SD, cfg, sched: SD, y_t, y_cond, view_count, angle, t, z, weighting=True, p="denoise_fn."):
    """p_mean_variance + p_sample, view_fusion.py:86-177.  Returns (y_prev, eps_hat, logits, weights)."""
    level = sched["gammas"].gather(-1, t).view(-1, 1)                                     # :98
    x, a, l = stack_views(y_cond, y_t, view_count, angle, level)
    out = unet_forward(sd, cfg, x, a, l, p=p)
    eps, logits, w = compose(out, view_count, weighting)
    return ddpm_update(sched, y_t, eps, t, z), eps, logits, w


def generate(sd: SD, cfg, sched: SD, y_cond, view_count, angle, y_T, zs: Sequence[Tensor],
             steps: Optional[Sequence[int]] = None, sample_num: int = 8, weighting=True, p="denoise_fn."):
    """generate view_fusion.py:179-214 with injected noise.

    `steps` defaults to reversed(range(T)); a shorter list gives a partial trajectory for
    tests (zs[j] is the draw used at steps[j]).  Returns (y_t, ret_arr, logit_arr, weight_arr, last).
    """
    T = sched["gammas"].shape[0]
    assert T > sample_num, "num_timesteps must greater than sample_num"
    inter = T // sample_num
    steps = list(reversed(range(T))) if steps is None else list(steps)
    y_t = y_T
    ret, logit_arr, weight_arr = [y_t], [], []
    B = y_cond.shape[0]
    for j, i in enumerate(steps):
        t = torch.full((B,), i, dtype=torch.long)
        y_t, _, logits, w = p_sample(sd, cfg, sched, y_t, y_cond, view_count, angle, t, zs[j], weighting, p)
        if i % inter == 0:
            ret.append(y_t); logit_arr.append(logits); weight_arr.append(w)
    ret = torch.stack(ret, dim=1)
    if weighting and logit_arr:
        logit_arr = torch.stack(logit_arr, dim=1)
        weight_arr = torch.stack(weight_arr, dim=1)
    return y_t, ret, logit_arr, weight_arr, ret[:, -1]


def train_gammas(sched: SD, t: Tensor, u: Tensor) -> Tensor:
    """view_fusion.py:232-237: gamma~ = (gamma[t] - gamma[t-1]) * u + gamma[t-1], shape (B,1)."""
    g = sched["gammas"]
    g1 = g.gather(-1, t - 1).view(-1, 1)
    g2 = g.gather(-1, t).view(-1, 1)
    return (g2 - g1) * u + g1


def train_loss(sd: SD, cfg, sched: SD, y_0, y_cond, view_count, angle, t, u, noise, weighting=True,
               p="denoise_fn."):
    """Training branch of ViewFusion.forward view_fusion.py:229-300 with the three random draws
    (t :231, u :234, noise :239) injected.  Returns (loss, eps_hat)."""
    sg = train_gammas(sched, t, u)
    y_noisy = sg.view(-1, 1, 1, 1).sqrt() * y_0 + (1 - sg.view(-1, 1, 1, 1)).sqrt() * noise   # :164
    x, a, l = stack_views(y_cond, y_noisy, view_count, angle, sg)
    out = unet_forward(sd, cfg, x, a, l, p=p)
    eps, _, _ = compose(out, view_count, weighting)
    return F.mse_loss(noise, eps), eps


def psnr(a: Tensor, b: Tensor) -> Tensor:
    """utils/metrics.py:6-8 (data range 1)."""
    mse = torch.mean((a - b) ** 2, dim=(1, 2, 3))
    return 20 * torch.log10(1.0 / torch.sqrt(mse))


def ssim(a: Tensor, b: Tensor, data_range: float = 1.0, win_size: int = 11, win_sigma: float = 1.5) -> Tensor:
    """utils/metrics.py:11-12 = pytorch_msssim.ssim(a, b, data_range=1.0, size_average=False) -> (B,).

    pytorch-msssim is a third-party dependency that is absent from this image (pinned ==1.0.0 in the reference's
    environment.yml:196), so this restates its published algorithm — PARITY UNPINNED for this function: 1-D Gaussian window
    (size 11, sigma 1.5, normalised to sum 1), applied separably along H then W as a VALID depth-wise convolution (no
    padding) to x, y, x*x, y*y, x*y; C1 = (0.01 L)^2, C2 = (0.03 L)^2;
    cs = (2 s_xy + C2) / (s_xx + s_yy + C2), ssim_map = (2 mu_x mu_y + C1) / (mu_x^2 + mu_y^2 + C1) * cs;
    mean of the map per (image, channel), then mean over channels."""
    C = a.shape[1]
    coords = torch.arange(win_size, dtype=torch.float32) - win_size // 2
    g = torch.exp(-(coords ** 2) / (2 * win_sigma ** 2))
    g = (g / g.sum()).to(a.dtype)

    def filt(x):
        x = F.conv2d(x, g.view(1, 1, -1, 1).repeat(C, 1, 1, 1), groups=C)      # along H
        return F.conv2d(x, g.view(1, 1, 1, -1).repeat(C, 1, 1, 1), groups=C)   # along W

    K1, K2 = 0.01, 0.03
    C1, C2 = (K1 * data_range) ** 2, (K2 * data_range) ** 2
    mu1, mu2 = filt(a), filt(b)
    s11, s22, s12 = filt(a * a) - mu1 * mu1, filt(b * b) - mu2 * mu2, filt(a * b) - mu1 * mu2
    cs = (2 * s12 + C2) / (s11 + s22 + C2)
    m = (2 * mu1 * mu2 + C1) / (mu1 * mu1 + mu2 * mu2 + C1) * cs
    return m.flatten(2).mean(-1).mean(1)


def process_batch_u8(views_u8: np.ndarray, perm: np.ndarray) -> Dict[str, np.ndarray]:
    """data/nmr_dataset.py:10-52 (`process_sample`) for a batch, with the view permutation given instead of drawn:
    views_u8 (B, V, H, W, C) uint8 as decoded from the dataset (webdataset's "rgb" decoder = uint8 / 255 in float32),
    perm (B, V) = the shuffled `images_idx`.  Returns target (B,C,H,W), cond (B,V-1,C,H,W), angle (B,1), all float32."""
    B, V = perm.shape
    images = views_u8.astype(np.float32) / np.float32(255.0)                 # "rgb" decode
    images = np.transpose(images, (0, 1, 4, 2, 3))                           # v h w c -> v c h w        (:15)
    shuffled = np.stack([images[b, perm[b]] for b in range(B)])              # cond_images = images[images_idx]  (:17-18)
    angle = (2 * np.pi / V * perm[:, :1]).astype(np.float32)                 # (:20-24)
    return {"target": shuffled[:, 0], "cond": shuffled[:, 1:], "angle": angle}


# --------------------------------------------------------------------------------------
# synthetic NMR-shaped inputs
# --------------------------------------------------------------------------------------

def synthetic_batch(B: int, N: int, size: int = 64, seed: int = 1234, ragged: bool = False, nmax: Optional[int] = None):
    """Deterministic numpy-PCG64 inputs: y_cond~U[0,1] (B,Nmax,3,S,S), y_0~U[0,1], angle=2*pi*k/24 (B,1),
    view_count (B,) int64 (all N, or U{1..N} when ragged), eps / y_T ~ N(0,1)."""
    rng = np.random.default_rng(seed)
    nmax = N if nmax is None else nmax
    f = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    y_cond = f(rng.uniform(0, 1, size=(B, nmax, 3, size, size)))
    y_0 = f(rng.uniform(0, 1, size=(B, 3, size, size)))
    angle = f((2 * math.pi / 24) * rng.integers(0, 24, size=(B, 1)))
    vc = rng.integers(1, N + 1, size=(B,)) if ragged else np.full((B,), N)
    noise = f(rng.standard_normal(size=(B, 3, size, size)))
    return dict(y_cond=y_cond, y_0=y_0, angle=angle, view_count=torch.from_numpy(vc.astype(np.int64)), noise=noise)


def normal_draws(n: int, shape, seed: int) -> List[Tensor]:
    rng = np.random.default_rng(seed)
    return [torch.from_numpy(rng.standard_normal(size=shape).astype(np.float32)) for _ in range(n)]

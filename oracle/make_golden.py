"""Pin the oracle against the REAL reference and write the golden fixtures.

TEST INFRASTRUCTURE.  Runs only in the build container, where the reference
checkout is mounted read-only at /root/reference (it does not exist on the GPU
box).  It

  1. imports `/root/reference/model/{unet,view_fusion}.py` UNMODIFIED,
  2. loads the deterministic state_dict of `vf_oracle.init_state_dict` (strict),
  3. runs the reference on seeded inputs, injecting randomness only through
     `torch.manual_seed` immediately before the reference call (the reference's
     draw order is re-played to recover t / u / z),
  4. asserts `oracle/vf_oracle.py` agrees with the reference, and
  5. writes the REFERENCE outputs to tests/golden/*.npz.

Usage:  PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")

import vf_oracle as O  # noqa: E402
from model.unet import UNet as RefUNet  # noqa: E402
from model.view_fusion import ViewFusion as RefViewFusion  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)
torch.set_num_threads(os.cpu_count() or 1)

BETA = {"train": dict(O.BETA_TRAIN), "test": dict(schedule="linear", num_timesteps=1000, linear_start=1e-4, linear_end=0.09)}


def rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def build_ref(cfg, seed, weighting=True):
    net = RefUNet(**{k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.items()})
    sd = O.init_state_dict(cfg, seed)
    net.load_state_dict(sd, strict=True)                     # same keys/shapes as Appendix C
    with contextlib.redirect_stdout(io.StringIO()):
        vf = RefViewFusion(net, BETA, weighting_train=weighting, weighting_inference=weighting)
    vf.set_new_noise_schedule(device=torch.device("cpu"), phase="train")
    full = {"denoise_fn." + k: v for k, v in sd.items()}
    return net, vf, sd, full


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = v
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"  wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


def golden_schedule():
    _, vf, _, _ = build_ref(O.TINY, 0)
    sched = O.make_schedule(**O.BETA_TRAIN)
    for k, v in sched.items():
        assert torch.equal(v, getattr(vf, k)), k             # bit-exact: same float64 numpy math
    save("schedule_train", **sched)
    return sched


def golden_unet(tag, cfg, R, seed):
    net, _, sd, _ = build_ref(cfg, seed)
    rng = np.random.default_rng(100 + seed)
    S = cfg["image_size"]
    x = torch.from_numpy(rng.standard_normal((R, cfg["in_channel"], S, S)).astype(np.float32))
    angle = torch.from_numpy((2 * np.pi / 24 * rng.integers(0, 24, (R, 1))).astype(np.float32))
    time = torch.from_numpy(rng.uniform(1e-4, 1.0, (R, 1)).astype(np.float32))
    with torch.no_grad():
        ref = net(x, angle, time)
        taps = {}
        mine = O.unet_forward(sd, cfg, x, angle, time, taps=taps)
    r = rel(mine, ref)
    print(f"unet[{tag}] oracle-vs-reference rel-L2 = {r:.2e}")
    assert r < 2e-6, r
    save(f"unet_{tag}", x=x, angle=angle, time=time, out=ref, seed=np.int64(seed),
         w_checksum=np.float64(sum(float(v.double().sum()) for v in sd.values())))


class _Fixed(torch.nn.Module):
    """Stand-in denoiser that returns a fixed tensor (lets the reference's composition/DDPM code run alone)."""

    def __init__(self, out):
        super().__init__()
        self.out = out

    def forward(self, x, angle, time):
        assert x.shape[0] == self.out.shape[0]
        return self.out


def golden_compose(sched):
    rng = np.random.default_rng(7)
    for tag, vc, weighting in (("ragged", [1, 4, 2, 6], True), ("full6", [6, 6, 6], True), ("mean", [2, 5, 1], False)):
        B, S = len(vc), 16
        R = sum(vc)
        cout = 6 if weighting else 3
        out = torch.from_numpy((2.0 * rng.standard_normal((R, cout, S, S))).astype(np.float32))
        y_t = torch.from_numpy(rng.standard_normal((B, 3, S, S)).astype(np.float32))
        y_cond = torch.zeros(B, max(vc), 3, S, S)
        angle = torch.zeros(B, 1)
        view_count = torch.tensor(vc, dtype=torch.long)
        with contextlib.redirect_stdout(io.StringIO()):
            vf = RefViewFusion(_Fixed(out), BETA, weighting_train=weighting, weighting_inference=weighting)
        vf.set_new_noise_schedule(device=torch.device("cpu"), phase="train")
        res = {}
        for tname, tv in (("hi", 1999), ("mid", 731), ("one", 1), ("zero", 0)):
            t = torch.full((B,), tv, dtype=torch.long)
            torch.manual_seed(55)
            y_prev, logits, w = vf.p_sample(y_t.clone(), y_cond, view_count, angle, t)
            torch.manual_seed(55)
            z = torch.randn_like(y_t)
            eps, lg, ww = O.compose(out, view_count, weighting)
            mine = O.ddpm_update(sched, y_t, eps, t, z)
            assert rel(mine, y_prev) < 1e-6, (tag, tname, rel(mine, y_prev))
            if weighting:
                assert torch.equal(lg, logits) and rel(ww, w) < 1e-6
            res[f"y_prev_{tname}"] = y_prev
            res[f"t_{tname}"] = t
        save(f"compose_{tag}", out=out, y_t=y_t, view_count=view_count, z=z, eps=eps,
             weights=(w if weighting else np.zeros(0, np.float32)), **res)


def golden_psample(tag, cfg, B, N, seed, steps, ragged):
    _, vf, sd, full = build_ref(cfg, seed)
    sched = O.make_schedule(**O.BETA_TRAIN)
    S = cfg["image_size"]
    bt = O.synthetic_batch(B, N, S, seed=1234 + seed, ragged=ragged)
    y_T = O.normal_draws(1, (B, 3, S, S), seed=99)[0]
    y_ref = y_T.clone()
    y_mine = y_T.clone()
    zs, ys, epss = [], [], []
    for j, i in enumerate(steps):
        t = torch.full((B,), i, dtype=torch.long)
        torch.manual_seed(1000 + j)
        y_ref, logits, w = vf.p_sample(y_ref, bt["y_cond"], bt["view_count"], bt["angle"], t)
        torch.manual_seed(1000 + j)
        z = torch.randn_like(y_ref)
        with torch.no_grad():
            y_mine, eps, lg, ww = O.p_sample(full, cfg, sched, y_mine, bt["y_cond"], bt["view_count"], bt["angle"], t, z)
        r = rel(y_mine, y_ref)
        assert r < 5e-6, (tag, i, r)
        assert rel(ww, w) < 5e-6
        zs.append(z); ys.append(y_ref.clone()); epss.append(eps)
    print(f"psample[{tag}] {len(steps)} steps, final oracle-vs-reference rel-L2 = {r:.2e}")
    save(f"psample_{tag}", y_cond=bt["y_cond"], angle=bt["angle"], view_count=bt["view_count"], y_T=y_T,
         steps=np.asarray(steps, np.int64), z=torch.stack(zs), y=torch.stack(ys), eps_oracle=torch.stack(epss),
         weights_last=w, logits_last=logits, seed=np.int64(seed))


def golden_train(tag, cfg, B, N, seed, ragged):
    net, vf, sd, full = build_ref(cfg, seed)
    sched = O.make_schedule(**O.BETA_TRAIN)
    S = cfg["image_size"]
    bt = O.synthetic_batch(B, N, S, seed=4321 + seed, ragged=ragged)
    T = sched["gammas"].shape[0]
    # replay the reference's draw order: randint (view_fusion.py:231) then rand (:234); noise is passed in.
    torch.manual_seed(77)
    t = torch.randint(1, T, (B,)).long()
    u = torch.rand((B, 1))
    torch.manual_seed(77)
    vf.zero_grad()
    loss = vf(y_cond=bt["y_cond"], view_count=bt["view_count"], angle=bt["angle"], y_0=bt["y_0"], noise=bt["noise"])
    loss.backward()
    req = {k: v.clone().requires_grad_(True) for k, v in full.items()}
    mine, eps = O.train_loss(req, cfg, sched, bt["y_0"], bt["y_cond"], bt["view_count"], bt["angle"], t, u, bt["noise"])
    mine.backward()
    assert abs(float(mine.detach()) - float(loss.detach())) < 1e-6 * max(1.0, abs(float(loss))), (float(mine.detach()), float(loss.detach()))
    gn = {k: float(p.grad.norm()) for k, p in net.named_parameters()}
    # biases that feed a 1-channel-per-group GroupNorm have a mathematically zero gradient (pure rounding noise
    # ~1e-8 in both implementations): they are compared against an absolute floor instead of their own norm.
    zero_floor = 1e-6 * max(gn.values())
    worst = 0.0
    for k, p in net.named_parameters():
        g_ref = p.grad
        g_mine = req["denoise_fn." + k].grad
        if gn[k] < zero_floor:
            assert float(g_mine.norm()) < zero_floor, k
            continue
        worst = max(worst, float((g_mine - g_ref).norm()) / gn[k])
    print(f"train[{tag}] loss {float(loss):.6f}; worst per-tensor grad rel-L2 oracle-vs-reference = {worst:.2e}")
    assert worst < 1e-4, worst
    names = list(gn)
    keep = [names[0], names[1], names[2], names[3], names[-1], names[-2]] + [n for n in names if "mid.0.attn" in n] \
        + [n for n in names if n.startswith("ups.1.res_block")]
    # kept tensors are stored as a strided subsample (<= 8192 values) of the flattened gradient; every tensor's
    # full L2 norm is stored in grad_norms.
    grads = {}
    for k in dict.fromkeys(keep):
        g = dict(net.named_parameters())[k].grad.reshape(-1)
        stride = max(1, (g.numel() + 8191) // 8192)
        grads["grad:" + k] = g[::stride]
    save(f"train_{tag}", y_cond=bt["y_cond"], y_0=bt["y_0"], angle=bt["angle"], view_count=bt["view_count"],
         noise=bt["noise"], t=t, u=u, loss=loss.detach(), eps=eps.detach(),
         grad_names=np.asarray(names), grad_norms=np.asarray([gn[n] for n in names], np.float64),
         seed=np.int64(seed), **grads)


def main():
    sched = golden_schedule()
    golden_unet("tiny", O.TINY, R=3, seed=0)
    golden_unet("small", O.SMALL_V100, R=2, seed=0)
    golden_compose(sched)
    golden_psample("tiny_ragged", O.TINY, B=3, N=4, seed=1, steps=[1999, 1998, 1000, 2, 1, 0], ragged=True)
    golden_psample("small_n3", O.SMALL_V100, B=2, N=3, seed=0, steps=[1999, 1998, 1, 0], ragged=False)
    golden_train("tiny_ragged", O.TINY, B=3, N=4, seed=2, ragged=True)
    golden_train("small_n3", O.SMALL_V100, B=2, N=3, seed=0, ragged=False)
    print("all oracle-vs-reference checks passed")


if __name__ == "__main__":
    main()

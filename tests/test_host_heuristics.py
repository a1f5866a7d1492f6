"""Host-side launch heuristics of the GroupNorm kernels through the C ABI's host-only test hooks (no device needed).

Small layers are scheduled by a latency model, time = waves x (prologue + iterations): every wave pays the statistics
prologue again, so a layer whose per-CTA streaming time is comparable to it must not be cut into many short CTAs
(ncu showed the 16x16 GroupNorm backward passes running in 4 waves)."""
import ctypes as C
import math

import pytest

from view_fusion_b200 import _lib

BF16 = _lib.VF_BF16
SMS = 148


def _fwd(images, S, Cc):
    th = C.c_int(0)
    s = _lib.load().vf_debug_gn_splits(images, S, S, Cc, BF16, C.byref(th))
    return s, th.value


def _bwd(images, S, Cc):
    return _lib.load().vf_debug_gn_bwd_splits(images, S, S, Cc, BF16)


@pytest.mark.parametrize("S,Cc", [(16, 192), (8, 320), (8, 640), (16, 384), (32, 128), (16, 128)])
def test_small_layers_run_in_one_wave_at_benchmark_size(S, Cc):
    images = 168                                        # B = 28, N = 6
    s, threads = _fwd(images, S, Cc)
    assert s >= 1 and threads > 0 and threads <= 256 and threads % (Cc // 8) == 0
    assert s * images <= SMS * 4, "forward: at most one wave of 4 resident CTAs per SM"
    sb = _bwd(images, S, Cc)
    assert sb >= 1
    waves = math.ceil(sb * images / (SMS * 2))
    assert waves <= 2, "backward: the reduce pass keeps 2 CTAs per SM resident"
    assert sb <= 3


def test_few_images_are_split_for_parallelism():
    # autoregressive sampling: B = 1, up to 24 views -> few images; one wave holds every split, so more splits only help
    s1, _ = _fwd(6, 16, 192)
    s2, _ = _fwd(168, 16, 192)
    assert s1 >= s2 and s1 >= 3
    assert _bwd(6, 16, 192) >= _bwd(168, 16, 192)


def test_large_layers_keep_the_wave_filling_rule():
    for S, Cc in [(64, 128), (64, 192), (32, 320)]:
        s, _ = _fwd(168, S, Cc)
        assert s >= 4, (S, Cc, s)                       # bandwidth-bound: enough CTAs to fill the machine several times
        assert _bwd(168, S, Cc) >= 3


def _slab(Cc, S, dtype=BF16):
    th, sm = C.c_int(0), C.c_int(0)
    sc = _lib.load().vf_debug_gn_bwd_slab(Cc, 32, S, S, dtype, C.byref(th), C.byref(sm))
    return sc, th.value, sm.value


@pytest.mark.parametrize("S,Cc", [(8, 192), (8, 320), (8, 512), (8, 640), (16, 128), (16, 192), (16, 320), (16, 384), (16, 512), (32, 64), (32, 128),
                                  (32, 256)])
def test_one_pass_groupnorm_backward_takes_the_layers_whose_slab_fits(S, Cc):
    """(image, slab of whole groups) resident in shared memory: every layer up to 32x32 x 256 channels of the benchmark UNet."""
    sc, threads, smem = _slab(Cc, S)
    gs = Cc // 32
    assert sc > 0 and Cc % sc == 0 and sc % gs == 0 and sc % 8 == 0, (sc, gs)          # whole groups, whole 16-byte vectors, covers C exactly
    assert sc * 2 >= 32, "rows of at least one 32-byte sector"
    assert 0 < threads <= 512 and threads % (sc // 8) == 0
    assert smem <= 90 * 1024 and smem >= 2 * S * S * sc * 2                             # x and dz of the slab live in shared memory


@pytest.mark.parametrize("S,Cc", [(64, 64), (64, 128), (64, 192), (32, 320), (32, 192)])
def test_one_pass_groupnorm_backward_leaves_the_large_layers_to_two_passes(S, Cc):
    assert _slab(Cc, S)[0] == 0


def test_one_pass_groupnorm_backward_fp32_and_bad_shapes():
    sc, threads, smem = _slab(192, 16, _lib.VF_F32)
    assert sc > 0 and sc % 6 == 0 and sc % 4 == 0 and smem <= 90 * 1024
    assert _lib.load().vf_debug_gn_bwd_slab(100, 32, 16, 16, BF16, None, None) < 0


def test_unsupported_shapes_are_rejected():
    assert _fwd(168, 16, 100)[0] < 0                    # channels not a multiple of the 16-byte vector
    assert _bwd(0, 16, 192) < 0


# ---- tcgen05 convolution: host-side plan (vf_debug_conv_tiling) over every layer shape of the benchmark UNet ----------
def _conv_plans(images=168):
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("tiling_table", os.path.join(os.path.dirname(__file__), "..", "scripts", "tiling_table.py"))
    tt = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tt)
    return [(name, S, segs, cout, kw, tt.plan(images, S, segs, cout, **kw)) for name, S, segs, cout, kw in tt.benchmark_layers()]


def test_conv_plans_respect_the_hardware_limits():
    for name, S, segs, cout, kw, p in _conv_plans():
        cout_pad = kw.get("cout_pad", cout)
        assert p["smem"] <= 227 * 1024, name
        assert p["tmem"] <= 512 and p["tmem"] >= 2 * p["G"] * p["bn"] and p["tmem"] & (p["tmem"] - 1) == 0, name   # two accumulator sets
        assert cout_pad % p["bn"] == 0 and p["bn"] % 16 == 0 and p["bn"] <= 256, name
        assert 1 <= p["G"] <= 4 and p["aS"] >= 2 and p["bS"] >= 1, name
        assert 1 <= p["grid"] <= SMS and p["grid"] == min(p["items"], SMS), name                                  # persistent CTAs, one per SM
        assert p["imgs"] * p["bn"] <= 1024, name                                                                   # bias table fits the epilogue's registers
        assert p["K"] == sum(c * k * k for c, k in segs), name
        assert p["items"] == math.ceil(p["rows"] / (128 * p["G"])) * (cout_pad // p["bn"]), name


def test_conv_modes_follow_the_layer_kind():
    plans = {name: p for name, *_, p in _conv_plans()}
    # every bf16-output layer leaves through the staged TMA epilogue; the fp32 final layer cannot
    assert all(p["epi"] == 1 for n, p in plans.items() if not n.startswith("final"))
    assert plans["final 64->6 (fp32 out)"]["epi"] == 0
    # PADDED -> FLAT 1x1 (qkv): the source is gathered as whole image lines, GEMM rows are the valid pixels only
    for n, W, rows in [("qkv 192 (inference)", 16, 168 * 256), ("qkv 192 (training)", 16, 168 * 256), ("qkv 320", 8, 168 * 64)]:
        assert plans[n]["aL"] == W and plans[n]["rows"] == rows and plans[n]["eL"] == 0
    # FLAT -> PADDED 1x1 (attention out-projection): 32-row tiles are whole lines of the padded output
    assert plans["attn out 192"]["eL"] == 16 and plans["attn out 320"]["eL"] == 8
    assert plans["conv0 (im2col 1x1, FLAT in)"]["eL"] == 0                     # W = 64: part of one line, plain 2D map
    # Downsample: output pixels only, one gathered A tile per (tap, 64 channels)
    for n, S, Cc in [("down 64", 64, 64), ("down 128", 32, 128), ("down 192", 16, 192)]:
        assert plans[n]["s2"] == Cc // 64 and plans[n]["rows"] == 168 * (S // 2) ** 2 and plans[n]["K"] == 9 * Cc
    # 3x3 layers run over PADDED rows
    assert plans["64->64 conv1"]["rows"] == 168 * 65 * 65 and plans["320->320 conv1"]["rows"] == 168 * 81


def test_conv_plan_scales_down_to_a_single_image():
    for name, S, segs, cout, kw, p in _conv_plans(images=1):
        assert p["items"] >= 1 and p["grid"] >= 1 and p["smem"] <= 227 * 1024, name

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


def test_unsupported_shapes_are_rejected():
    assert _fwd(168, 16, 100)[0] < 0                    # channels not a multiple of the 16-byte vector
    assert _bwd(0, 16, 192) < 0

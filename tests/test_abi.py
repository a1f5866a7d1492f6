"""CPU: the C-ABI library loads, and exports every symbol include/viewfusion_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "viewfusion_b200.h")


@pytest.fixture(scope="module")
def lib():
    from view_fusion_b200 import build, _lib
    build.build()                      # nvcc cross-compiles without a GPU; incremental
    return _lib.load()


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vf_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    from view_fusion_b200 import _lib
    decl = declared_symbols()
    assert len(decl) >= 25
    for s in decl:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert sorted(_lib.SYMBOLS) == decl, "view_fusion_b200/_lib.py binds a different symbol set than the header declares"


def test_probe_library_is_separate(lib):
    """The tcgen05 hardware probes ship in their own library: the product .so must not export them."""
    from view_fusion_b200 import _lib
    probes = _lib.load_probes()
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "viewfusion_b200_probes.h")).read(), flags=re.S)
    decl = sorted(set(re.findall(r"\b(vf_debug_[a-z0-9_]+)\s*\(", src)))
    assert decl == sorted(_lib.PROBE_SYMBOLS)
    for s in decl:
        assert hasattr(probes, s)
        assert not hasattr(lib, s), f"{s} must not be exported by the product library"


def test_abi_version_and_error_string(lib):
    assert lib.vf_abi_version() == 2
    assert isinstance(lib.vf_last_error(), bytes)


def test_plan_param_table_matches_reference_layout(lib):
    """Host-only: the plan's parameter table == reference state_dict order/shapes."""
    import vf_oracle as O
    from view_fusion_b200 import _lib
    for cfg in (O.SMALL_V100, O.TINY):
        c = _lib.UnetConfig()
        c.in_channel, c.out_channel, c.inner_channel, c.norm_groups = cfg["in_channel"], cfg["out_channel"], cfg["inner_channel"], 32
        c.n_mults = len(cfg["channel_mults"])
        for i, m in enumerate(cfg["channel_mults"]):
            c.channel_mults[i] = m
        c.n_attn_res = len(cfg["attn_res"])
        for i, r in enumerate(cfg["attn_res"]):
            c.attn_res[i] = r
        c.res_blocks, c.image_size = cfg["res_blocks"], cfg["image_size"]
        h = ctypes.c_void_p()
        assert lib.vf_unet_create(ctypes.byref(c), _lib.VF_F32, ctypes.byref(h)) == 0, lib.vf_last_error()
        want = O.param_shapes(cfg)
        assert lib.vf_unet_num_params(h) == len(want)
        buf = ctypes.create_string_buffer(256)
        shape = (ctypes.c_int64 * 4)()
        nd = ctypes.c_int()
        for i, (name, shp, _) in enumerate(want):
            assert lib.vf_unet_param_info(h, i, buf, 256, shape, ctypes.byref(nd)) == 0
            assert buf.value.decode() == name and tuple(shape[: nd.value]) == tuple(shp)
        assert lib.vf_unet_workspace_bytes(h, 4) > 0 and lib.vf_unet_packed_bytes(h) > 0
        lib.vf_unet_destroy(h)


def test_bad_config_is_an_error_not_a_crash(lib):
    from view_fusion_b200 import _lib
    c = _lib.UnetConfig()
    c.in_channel, c.out_channel, c.inner_channel, c.norm_groups, c.n_mults = 6, 6, 48, 32, 1
    c.channel_mults[0] = 1
    c.res_blocks, c.image_size = 1, 16
    h = ctypes.c_void_p()
    assert lib.vf_unet_create(ctypes.byref(c), _lib.VF_BF16, ctypes.byref(h)) == -1
    assert b"norm_groups" in lib.vf_last_error() or b"channels" in lib.vf_last_error()


def _declared_arity():
    """{function: number of parameters} parsed from the header's prototypes."""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(vf_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", src, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        out[name] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def test_ctypes_signatures_have_the_declared_arity(lib):
    """A binding with the wrong number of arguments would still 'work' through ctypes and corrupt the call."""
    arity = _declared_arity()
    checked = 0
    for name, n in arity.items():
        fn = getattr(lib, name)
        if fn.argtypes is None:
            continue
        assert len(fn.argtypes) == n, f"{name}: header declares {n} parameters, _lib.py binds {len(fn.argtypes)}"
        checked += 1
    assert checked >= 40

"""Host-only: the tcgen05 convolution's plan (vf_debug_conv_tiling) for every convolution of the benchmark UNet at B=28, N=6.
python scripts/tiling_table.py [images]      (no GPU needed)"""
import ctypes as C
import sys
sys.path.insert(0, ".")
from view_fusion_b200 import _lib

NAMES = ["bn", "G", "aS", "bS", "res", "smem", "tmem", "items", "grid", "epi", "aL", "eL", "s2", "rows", "imgs", "K"]


def plan(images, S, segs, cout, stride=1, in_padded=1, out_padded=1, qkv_split=0, cout_pad=None, out_f32=False, name=""):
    a = _lib.ConvArgs()
    a.dtype, a.images, a.H, a.W, a.n_seg = _lib.VF_BF16, images, S, S, len(segs)
    a.in_padded, a.out_padded, a.stride = in_padded, out_padded, stride
    for i, (c, k) in enumerate(segs):
        a.src_c[i], a.ksize[i] = c, k
    a.cout, a.cout_pad = cout, cout_pad or cout
    a.out_dtype, a.out_ld, a.qkv_split = (_lib.VF_F32 if out_f32 else _lib.VF_BF16), (8 if out_f32 else cout), qkv_split
    # planning only looks at WHETHER the epilogue has a bias / statistics to handle (never dereferenced): every layer of the UNet
    # except the qkv projections has a bias; every bf16 output except qkv feeds a GroupNorm (fused statistics)
    if "qkv" not in name:
        a.bias = 1
        if not out_f32:
            a.stats = 1
    out = (C.c_int * 16)()
    rc = _lib.load().vf_debug_conv_tiling(C.byref(a), out)
    if rc != 0:
        raise RuntimeError(_lib.load().vf_last_error().decode())
    return dict(zip(NAMES, out))


def benchmark_layers():
    """(name, S, segs, cout, kwargs) of the small-v100 UNet's 84 convolution launches, grouped by shape."""
    L = [("conv0 (im2col 1x1, FLAT in)", 64, [(64, 1)], 64, dict(in_padded=0)),
         ("64->64 conv1", 64, [(64, 3)], 64, {}), ("64->64 conv2 (+identity)", 64, [(64, 3), (64, 1)], 64, {}),
         ("down 64", 64, [(64, 3)], 64, dict(stride=2)),
         ("64->128 conv1", 32, [(64, 3)], 128, {}), ("128 conv2 (+res 64)", 32, [(128, 3), (64, 1)], 128, {}),
         ("128->128 conv1", 32, [(128, 3)], 128, {}), ("128 conv2 (+identity)", 32, [(128, 3), (128, 1)], 128, {}),
         ("down 128", 32, [(128, 3)], 128, dict(stride=2)),
         ("128->192 conv1", 16, [(128, 3)], 192, {}), ("192->192 conv1", 16, [(192, 3)], 192, {}),
         ("192 conv2 (+identity)", 16, [(192, 3), (192, 1)], 192, {}),
         ("qkv 192 (inference)", 16, [(192, 1)], 576, dict(out_padded=0)), ("qkv 192 (training)", 16, [(192, 1)], 576, dict(out_padded=0, qkv_split=192)),
         ("attn out 192", 16, [(192, 1)], 192, dict(in_padded=0)),
         ("down 192", 16, [(192, 3)], 192, dict(stride=2)),
         ("192->320 conv1", 8, [(192, 3)], 320, {}), ("320->320 conv1", 8, [(320, 3)], 320, {}),
         ("320 conv2 (+identity)", 8, [(320, 3), (320, 1)], 320, {}),
         ("qkv 320", 8, [(320, 1)], 960, dict(out_padded=0)), ("attn out 320", 8, [(320, 1)], 320, dict(in_padded=0)),
         ("640->320 conv1", 8, [(640, 3)], 320, {}), ("320 conv2 (+res 320+320)", 8, [(320, 3), (320, 1), (320, 1)], 320, {}),
         ("512->320 conv1", 8, [(512, 3)], 320, {}), ("up 320 @16", 16, [(320, 3)], 320, {}),
         ("512->192 conv1", 16, [(512, 3)], 192, {}), ("192 conv2 (+res 320+192)", 16, [(192, 3), (320, 1), (192, 1)], 192, {}),
         ("384->192 conv1", 16, [(384, 3)], 192, {}), ("320->192 conv1", 16, [(320, 3)], 192, {}),
         ("up 192 @32", 32, [(192, 3)], 192, {}), ("320->128 conv1", 32, [(320, 3)], 128, {}),
         ("128 conv2 (+res 192+128)", 32, [(128, 3), (192, 1), (128, 1)], 128, {}), ("256->128 conv1", 32, [(256, 3)], 128, {}),
         ("192->128 conv1", 32, [(192, 3)], 128, {}), ("up 128 @64", 64, [(128, 3)], 128, {}),
         ("192->64 conv1", 64, [(192, 3)], 64, {}), ("64 conv2 (+res 128+64)", 64, [(64, 3), (128, 1), (64, 1)], 64, {}),
         ("128->64 conv1", 64, [(128, 3)], 64, {}), ("64 conv2 (+res 64+64)", 64, [(64, 3), (64, 1), (64, 1)], 64, {}),
         ("final 64->6 (fp32 out)", 64, [(64, 3)], 6, dict(cout_pad=16, out_padded=0, out_f32=True))]
    return L


if __name__ == "__main__":
    images = int(sys.argv[1]) if len(sys.argv) > 1 else 168
    print(f"{'layer':34s} " + " ".join(f"{n:>6s}" for n in NAMES) + "   waves")
    for name, S, segs, cout, kw in benchmark_layers():
        p = plan(images, S, segs, cout, name=name, **kw)
        print(f"{name:34s} " + " ".join(f"{p[n]:6d}" for n in NAMES) + f"   {p['items'] / 148:5.2f}")

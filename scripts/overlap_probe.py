"""Would two half-batches on two streams overlap the HBM-bound GroupNorm passes of one with the tensor-bound convolutions of the
other?  One UNet forward over 168 view-images vs two forwards over 84 each (a) back to back on one stream, (b) on two streams.
python scripts/overlap_probe.py"""
import contextlib, io, sys
import torch
sys.path.insert(0, ".")
from view_fusion_b200 import UNet
from bench import SMALL

dev = "cuda"
with contextlib.redirect_stdout(io.StringIO()):
    nets = [UNet(**SMALL, precision="bf16").to(dev) for _ in range(2)]
S = 64


def inputs(unet, images, rows):
    x0 = torch.randn(images * S * S * unet.k0, device=dev).bfloat16()
    level = torch.rand(rows, device=dev)
    angle = torch.rand(rows, device=dev)
    img_row = (torch.arange(images, device=dev, dtype=torch.int32) * rows // images).int()
    out = torch.empty(images * S * S * 8, device=dev)
    return x0, images, level, angle, img_row, out


full = inputs(nets[0], 168, 28)
halves = [inputs(nets[i], 84, 14) for i in range(2)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]


def one():
    nets[0].run_packed(*full, stash=False)


def two_serial():
    for i in range(2):
        nets[i].run_packed(*halves[i], stash=False)


def two_streams():
    cur = torch.cuda.current_stream()
    for i in range(2):
        streams[i].wait_stream(cur)
        with torch.cuda.stream(streams[i]):
            nets[i].run_packed(*halves[i], stash=False)
    for i in range(2):
        cur.wait_stream(streams[i])


def graphed(fn):
    fn(); fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g.replay


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


with torch.no_grad():
    for name, fn in [("one forward, 168 view-images", one), ("two forwards of 84, one stream", two_serial), ("two forwards of 84, two streams", two_streams)]:
        te = timed(fn)
        tg = timed(graphed(fn))
        print(f"{name:34s}: eager {te:6.3f} ms   CUDA graph {tg:6.3f} ms", flush=True)

"""In-tree build of the C-ABI library `libviewfusion_b200.so` (nvcc, sm_100a only), plus
`libviewfusion_b200_probes.so` = the same objects + `csrc/k_debug.cu` (tcgen05 hardware probes used by tests and scripts; never
loaded by the product package).

    python -m view_fusion_b200.build            # incremental
    python -m view_fusion_b200.build --force

nvcc cross-compiles without a GPU, so this also is the CPU-side "does it build" check.  Objects are cached under
`view_fusion_b200/csrc/_build/` keyed by a hash of the source, the headers and the flags.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.environ.get("VF_B200_OUT") or os.path.join(HERE, "libviewfusion_b200.so")      # A/B builds: VF_B200_OUT + VF_NVCC_EXTRA
OUT_PROBES = os.path.join(HERE, "libviewfusion_b200_probes.so")    # product objects + hardware probes: tests / scripts only
PROBE_SOURCES = ("k_debug.cu",)
BUILD = os.path.join(CSRC, os.environ.get("VF_B200_BUILD_DIR", "_build"))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
] + os.environ.get("VF_NVCC_EXTRA", "").split()


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cuh", ".h")):
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(src, hdr, force, verbose):
    path = os.path.join(CSRC, src)
    with open(path, "rb") as fh:
        key = hashlib.sha256(fh.read() + hdr.encode()).hexdigest()[:16]
    obj = os.path.join(BUILD, f"{src[:-3]}.{key}.o")
    if os.path.exists(obj) and not force:
        return obj, False
    for old in os.listdir(BUILD):
        if old.startswith(src[:-3] + ".") and old.endswith(".o"):
            os.unlink(os.path.join(BUILD, old))
    cmd = [NVCC, *FLAGS, "-c", path, "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(r.stderr)
    return obj, True


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    hdr = _headers_digest()
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        res = list(ex.map(lambda s: _compile(s, hdr, force, verbose), _sources()))
    srcs = _sources()
    objs = [o for (o, _), src in zip(res, srcs) if src not in PROBE_SOURCES]          # the product library carries no probes
    objs_all = [o for o, _ in res]
    changed = any(c for _, c in res)
    targets = ((OUT, objs),) if os.environ.get("VF_B200_OUT") else ((OUT, objs), (OUT_PROBES, objs_all))     # A/B builds: product library only
    for out, ob in targets:
        if changed or not os.path.exists(out) or force:
            cmd = [NVCC, "-shared", "-o", out, *ob, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

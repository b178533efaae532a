"""Builds libidprimer_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

One object per translation unit, compiled in parallel (the kernel families are independent), then one link."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libidprimer_b200.so")
SOURCES = ["idp_panel.cpp", "idp_bam.cpp", "idp_ctx.cu", "idp_scan.cu", "idp_sw.cu"]
HEADERS = ["idp_internal.h", "idp_device.cuh", "idp_launch.h", "idp_kernels.cuh", "idp_scan.cuh", os.path.join("..", "..", "include", "idprimer.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newest_header() -> float:
    return max(os.path.getmtime(os.path.join(CSRC, f)) for f in HEADERS)


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES) or _newest_header() > t


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    hdr = _newest_header()

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        path = os.path.join(CSRC, src)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(path), hdr):
            return obj
        cmd = [nvcc, "-O3", "-std=c++17", *ARCH, "-lineinfo", "-Xcompiler", "-fPIC", "-c", "-o", obj, path]
        if verbose and src.endswith(".cu"):
            cmd += ["-Xptxas", "-v"]
        subprocess.check_call(cmd)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    subprocess.check_call([nvcc, *ARCH, "-shared", "-Xcompiler", "-fPIC", "-o", LIB, *objs])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

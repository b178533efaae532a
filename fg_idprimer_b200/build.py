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
HEADERS = ["idp_internal.h", "idp_device.cuh", "idp_launch.h", "idp_kernels.cuh", "idp_scan.cuh", "idp_scanfx.cuh", os.path.join("..", "..", "include", "idprimer.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newest_header() -> float:
    return max(os.path.getmtime(os.path.join(CSRC, f)) for f in HEADERS)


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES) or _newest_header() > t


# the alignment choices nothing in the reference pins, all flipped (idp_internal.h, AlignPolicy): the build tests/test_align_policy.py
# keeps parity-green against the oracle compiled with the same flags
ALT_POLICY_FLAGS = ["-DIDP_POLICY_LEFT_FROM_UP=1", "-DIDP_POLICY_UP_FROM_LEFT=1", "-DIDP_POLICY_EXTEND_OVER_OPEN=1", "-DIDP_POLICY_TIE_UP_BEFORE_LEFT=1"]
ALT_LIB = os.path.join(HERE, "libidprimer_b200_altpolicy.so")


def build_alt_policy(force: bool = False) -> str:
    """libidprimer_b200_altpolicy.so: the same sources with every AlignPolicy switch flipped."""
    return build(force=force, lib=ALT_LIB, obj_dir=OBJ + "_altpolicy", extra=ALT_POLICY_FLAGS)


def build(force: bool = False, verbose: bool = False, lib: str = LIB, obj_dir: str = OBJ, extra=()) -> str:
    def is_stale() -> bool:
        if not os.path.exists(lib):
            return True
        t = os.path.getmtime(lib)
        return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES) or _newest_header() > t
    if not force and not is_stale():
        return lib
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(obj_dir, exist_ok=True)
    hdr = _newest_header()

    def compile_one(src: str) -> str:
        obj = os.path.join(obj_dir, os.path.splitext(src)[0] + ".o")
        path = os.path.join(CSRC, src)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(path), hdr):
            return obj
        cmd = [nvcc, "-O3", "-std=c++17", *ARCH, "-lineinfo", "-Xcompiler", "-fPIC", *extra, "-c", "-o", obj, path]
        if verbose and src.endswith(".cu"):
            cmd += ["-Xptxas", "-v"]
        subprocess.check_call(cmd)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    subprocess.check_call([nvcc, *ARCH, "-shared", "-Xcompiler", "-fPIC", "-o", lib, *objs])
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    if "--alt-policy" in sys.argv:
        print(build_alt_policy(force="--force" in sys.argv))

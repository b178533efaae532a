"""Builds and runs the C++ host mirror (include/idprimer.hpp) test program: host mode here, gpu mode on the B200."""
import os
import subprocess

import pytest

import helpers as H
from fg_idprimer_b200 import _lib as L

EXE = os.path.join(H.ROOT, "tests", "cpp", "mirror_test")


def _build():
    L.lib()  # the shared library must exist
    src = os.path.join(H.ROOT, "tests", "cpp", "mirror_test.cpp")
    libdir = os.path.join(H.ROOT, "fg_idprimer_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(H.ROOT, "include"), src, "-o", EXE,
                           "-L", libdir, "-l:libidprimer_b200.so", f"-Wl,-rpath,{libdir}"])


def test_cpp_mirror_host():
    _build()
    out = subprocess.run([EXE, "host"], capture_output=True, text=True)
    assert out.returncode == 0 and "OK host" in out.stdout, out.stderr


@pytest.mark.gpu
def test_cpp_mirror_gpu_known_answers():
    _build()
    out = subprocess.run([EXE, "gpu"], capture_output=True, text=True)
    assert out.returncode == 0 and "OK gpu" in out.stdout, out.stderr

"""Builds and runs the C++ facade test (include/m3tsz_b200.hpp over the C ABI)."""
import os
import subprocess

import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
SRC = os.path.join(ROOT, "tests", "cpp", "test_facade.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "test_facade.bin")


def _build():
    libdir = os.path.join(ROOT, "m3_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), SRC, "-o", EXE,
           "-L", libdir, "-lm3tsz_b200", "-Wl,-rpath," + libdir, "-L/usr/local/cuda/lib64",
           "-Wl,-rpath,/usr/local/cuda/lib64"]
    subprocess.check_call(cmd)


def test_cpp_facade_builds_and_refuses_without_gpu():
    import torch
    _build()
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu-marked test")
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 3 and "NO_DEVICE" in r.stdout, (r.returncode, r.stdout, r.stderr)


@pytest.mark.gpu
def test_cpp_facade_goldens_on_gpu():
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 0 and "PASS" in r.stdout, (r.returncode, r.stdout, r.stderr)

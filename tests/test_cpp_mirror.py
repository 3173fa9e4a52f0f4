"""The C++ header mirror (include/nvbio_b200/nvbio_b200.hpp) compiles with plain g++ against the C ABI (CPU),
and reproduces the reference-asserted banded problem on the GPU."""
import os
import subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "abi_example")


def build():
    from nvbio_b200 import build as b
    b.build()
    cmd = ["g++", "-O1", "-std=c++14", "-I" + os.path.join(ROOT, "include"), "-I/usr/local/cuda/include",
           os.path.join(ROOT, "tests", "cpp", "abi_example.cpp"), "-o", EXE,
           "-L" + os.path.join(ROOT, "nvbio_b200"), "-lnvbio_b200", "-L/usr/local/cuda/lib64", "-lcudart",
           "-Wl,-rpath," + os.path.join(ROOT, "nvbio_b200"), "-Wl,-rpath,/usr/local/cuda/lib64"]
    subprocess.check_call(cmd)


def test_cpp_mirror_compiles_and_links():
    build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_mirror_runs():
    build()
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "score -11 sink (165,150)" in r.stdout

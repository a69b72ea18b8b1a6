"""The header-only C++ mirror (include/concrete_ntt.hpp): compiles and links against libcntt_b200.so on the
CPU box; runs the reference's README / examples flow on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "facade_test")


def build_exe():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    libdir = os.path.join(ROOT, "concrete-ntt_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "facade_test.cpp"), "-o", EXE,
                           "-L", libdir, "-l:libcntt_b200.so", "-Wl,-rpath," + libdir])
    return EXE


def test_facade_compiles_and_links():
    assert os.path.exists(build_exe())


@pytest.mark.gpu
def test_facade_runs_reference_examples():
    exe = build_exe()
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "facade ok" in out.stdout

"""Compiles tests/cpp/test_classes.cpp (the reference's gtest cases against this repo's NVStrings/NVCategory/NVText
C++ classes) with g++ and runs it on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_exe():
    out_dir = os.path.join(ROOT, "tests", "cpp", "_build")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, "test_classes")
    lib_dir = os.path.join(ROOT, "custrings_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_classes.cpp"), "-o", exe,
           "-L" + lib_dir, "-lcustr", "-Wl,-rpath," + lib_dir]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def test_cpp_classes_compile_and_link():
    """no GPU needed: the class surface compiles and links against libcustr.so"""
    from custrings_b200 import build
    build.build()
    assert os.path.exists(build_exe())


@pytest.mark.gpu
def test_cpp_classes_run():
    r = subprocess.run([build_exe()], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout + r.stderr

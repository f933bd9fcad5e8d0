"""Builds tests/sim/_build/libcustr_sim.so: the host simulation harness (sim.cu) linked against libcustr.so."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def build():
    from custrings_b200 import build as b
    lib = b.build()
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "libcustr_sim.so")
    src = os.path.join(HERE, "sim.cu")
    deps = [src, lib] + b._headers()
    if not b._stale(out, deps):
        return out
    cmd = [b.NVCC] + b.ARCH + ["-O2", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-shared", src, "-o", out,
                               "-L" + os.path.dirname(lib), "-lcustr", "-Xlinker", "-rpath," + os.path.dirname(lib)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("sim build failed:\n" + r.stdout + r.stderr)
    return out


if __name__ == "__main__":
    print(build())

"""Builds libcustr.so (hand-written sm_100a CUDA + host C++) in-tree with nvcc.  No torch involved."""
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libcustr.so")
NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-I", os.path.join(ROOT, "include")]

# (source, extra defines, object name); regex_item.cu is built once per chain-length group so its instantiations compile in parallel
SOURCES = [("regex_bits.cu", [], None), ("regex_item.cu", ["-DITEM_NS_GROUP=0"], "regex_item_g0"), ("regex_item.cu", ["-DITEM_NS_GROUP=1"], "regex_item_g1"),
           ("regex_item.cu", ["-DITEM_NS_GROUP=2"], "regex_item_g2"), ("regex_item.cu", ["-DITEM_NS_GROUP=3"], "regex_item_g3"),
           ("regex.cu", [], None), ("column.cu", [], None), ("attrs.cu", [], None), ("find.cu", [], None), ("split.cu", [], None), ("category.cu", [], None),
           ("regex_jit.cu", [], None), ("regex_bits_lower.cpp", [], None), ("regex_compile.cpp", [], None), ("classes.cpp", [], None)]
# kernel headers embedded into the library for the run-time compiled plan kernels (regex_jit.cu); "cstdint" / "cuda_runtime.h"
# are stand-ins: NVRTC has no host headers
JIT_HEADERS = ["device_utils.cuh", "regex_bits_plan.h", "regex_bits_dev.cuh", "regex_chain.cuh", "regex_chain64.cuh", "regex_chain_item.cuh"]
JIT_STANDINS = {
    "cstdint": "typedef signed char int8_t; typedef unsigned char uint8_t; typedef short int16_t; typedef unsigned short uint16_t; "
               "typedef int int32_t; typedef unsigned int uint32_t; typedef long long int64_t; typedef unsigned long long uint64_t; "
               "typedef unsigned long uintptr_t; typedef unsigned long size_t;\n",
    "cuda_runtime.h": "\n",
}


def _embed_jit_headers():
    out = os.path.join(OBJ, "jit_headers.inc")
    srcs = [os.path.join(CSRC, h) for h in JIT_HEADERS]
    if not _stale(out, srcs + [os.path.abspath(__file__)]):
        return
    parts = []
    for name, text in list(JIT_STANDINS.items()) + [(h, open(os.path.join(CSRC, h)).read()) for h in JIT_HEADERS]:
        chunks = [text[i:i + 8000] for i in range(0, len(text), 8000)] or [""]  # string literals have a length limit: concatenate
        lit = "\n".join('R"CUSTRJIT(%s)CUSTRJIT"' % c for c in chunks)
        parts.append('{"%s",\n%s},\n' % (name, lit))
    with open(out, "w") as f:
        f.write("".join(parts))



def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh", ".inc"))]
    hs.append(os.path.join(ROOT, "include", "custr.h"))
    return hs


def _compile(item, verbose):
    src, defines, name = item
    path = os.path.join(CSRC, src)
    obj = os.path.join(OBJ, (name or src.rsplit(".", 1)[0]) + ".o")
    if not _stale(obj, [path] + _headers()):
        return obj
    cmd = [NVCC] + ARCH + COMMON + defines + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    if verbose:
        sys.stderr.write(r.stderr)
    return obj


def build(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    flags_inc = os.path.join(CSRC, "unicode_flags.inc")
    gen = os.path.join(ROOT, "tools", "gen_unicode_flags.py")
    if force or _stale(flags_inc, [gen, os.path.join(ROOT, "tools", "unicode_delta.txt")]):
        subprocess.run([sys.executable, gen], check=True, capture_output=True)
    cases_inc = os.path.join(CSRC, "unicode_cases.inc")
    gen2 = os.path.join(ROOT, "tools", "gen_unicode_cases.py")
    if force or _stale(cases_inc, [gen, gen2, os.path.join(ROOT, "tools", "unicode_cases_delta.txt")]):
        subprocess.run([sys.executable, gen2], check=True, capture_output=True)
    if force:
        shutil.rmtree(OBJ)
        os.makedirs(OBJ)
    _embed_jit_headers()
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s[0]))]
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), srcs))
    if _stale(LIB, objs):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    _build_pyni()
    return LIB


def _build_pyni():
    """CPython extension modules pyniNVStrings / pyniNVCategory / pyniNVText (custrings_b200/pyni/): the reference's binding
    layer for the hot path, over libcustr.so's C-ABI.  Plain g++ against this interpreter's Python.h."""
    import sysconfig
    inc = sysconfig.get_paths()["include"]
    if not os.path.exists(os.path.join(inc, "Python.h")):
        sys.stderr.write("custrings_b200.build: Python.h not found, pyni modules not built\n")
        return
    pdir = os.path.join(HERE, "pyni")
    cxx = shutil.which("g++") or "g++"
    for src, mod in (("pystrings.cpp", "pyniNVStrings"), ("pycategory.cpp", "pyniNVCategory"), ("pytext.cpp", "pyniNVText")):
        out = os.path.join(pdir, mod + ".so")
        deps = [os.path.join(pdir, src), os.path.join(pdir, "pyni_common.h"), os.path.join(ROOT, "include", "custr.h")]
        if not _stale(out, deps):
            continue
        cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-I", inc, os.path.join(pdir, src), "-o", out, "-L", HERE, "-lcustr",
               "-Wl,-rpath,$ORIGIN/.."]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("pyni build failed for %s:\n%s" % (src, r.stderr))


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))

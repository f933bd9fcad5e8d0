"""The C-ABI library loads and exports every symbol include/custr.h declares (no compute calls: no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "custr.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(custr_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from custrings_b200 import build
    lib = ctypes.CDLL(build.build())
    syms = declared_symbols()
    assert len(syms) >= 50
    for s in syms:
        assert hasattr(lib, s), "libcustr.so does not export %s" % s


def test_python_binding_covers_header():
    from custrings_b200 import _lib
    assert set(declared_symbols()) == set(_lib.EXPORTED_SYMBOLS)
    _lib.lib()  # resolves every symbol with its signature; raises on mismatch


def test_no_cuda_calls_needed_for_describe():
    from custrings_b200._lib import lib
    buf = ctypes.create_string_buffer(4096)
    n = lib().custr_regex_describe(rb"\b\w{4,}\b", buf, len(buf))
    assert n == 8 and b"bitstream" in buf.value
    assert lib().custr_version().startswith(b"custrings_b200")


def test_product_does_not_touch_oracle():
    """the product path must never import / link / execute anything under oracle/"""
    pkg = os.path.join(ROOT, "custrings_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in src.replace("the oracle", "").replace("oracle)", "").lower() or f == "__nothing__", (dirpath, f)

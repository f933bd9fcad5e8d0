#!/usr/bin/env python3
"""Builds an A/B variant of libcustr.so that differs only in the item-kernel objects: custrings_b200/_variants/libcustr_<name>.so
    python tools/build_variant.py NAME -DMACRO [-DMACRO ...]        run with  CUSTR_LIB=custrings_b200/_variants/libcustr_NAME.so
Only regex_item.cu group 1 (chains of 3-4 steps: the headline) is rebuilt with the macros; every other object is reused."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from custrings_b200 import build as b  # noqa: E402

name, macros = sys.argv[1], sys.argv[2:]
b.build()
out_dir = os.path.join(b.HERE, "_variants")
os.makedirs(out_dir, exist_ok=True)
obj = os.path.join(out_dir, "regex_item_g1_%s.o" % name)
subprocess.run([b.NVCC] + b.ARCH + b.COMMON + ["-DITEM_NS_GROUP=1"] + macros + ["-c", os.path.join(b.CSRC, "regex_item.cu"), "-o", obj], check=True)
objs = [os.path.join(b.OBJ, f) for f in sorted(os.listdir(b.OBJ)) if f.endswith(".o") and f != "regex_item_g1.o"] + [obj]
lib = os.path.join(out_dir, "libcustr_%s.so" % name)
subprocess.run([b.NVCC] + b.ARCH + ["-shared", "-o", lib] + objs + ["-lcudart"], check=True)
print(lib)

#!/usr/bin/env python3
"""CUSTR_TRACE=1 python tools/trace_replace.py — host-side phase times of replace_re / replace on the C2 column (development aid:
the library synchronises at every trace point, so the sum is larger than the untraced call)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from custrings_b200 import nvstrings  # noqa: E402
from custrings_b200.workloads import c2_corpus  # noqa: E402

n, nbytes = int(os.environ.get("ROWS", 10_000_000)), int(os.environ.get("BYTES", 1 << 30))
chars, offsets, validity, nulls = c2_corpus(n, nbytes)
col = nvstrings.from_offsets(chars, offsets, n, validity, nulls)
import time
import torch
for i in range(6):
    print("-- call %d: replace_re" % i, file=sys.stderr, flush=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a.record()
    r = col.replace(r"\b\w{4,}\b", "#")
    b.record()
    torch.cuda.synchronize()
    print("   events %.3f ms, wall %.3f ms" % (a.elapsed_time(b), (time.perf_counter() - t0) * 1e3), file=sys.stderr, flush=True)
    del r
for i in range(3):
    print("-- call %d: literal replace" % i, file=sys.stderr, flush=True)
    r = col.replace("ab", "X", regex=False)
    del r

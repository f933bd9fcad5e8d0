#!/usr/bin/env python3
"""contains_re call time against column size (rows of C2): where the per-call fixed cost sits.  One GPU."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from custrings_b200 import nvstrings  # noqa: E402
from custrings_b200._lib import lib  # noqa: E402
from custrings_b200.workloads import c2_corpus, slice_rows  # noqa: E402

n, nbytes = 10_000_000, 1 << 30
chars, offsets, validity, nulls = c2_corpus(n, nbytes)
L = lib()
pat = rb"\b\w{4,}\b"
for rows in (10_000_000, 5_000_000, 2_500_000, 1_250_000, 312_500, 10_000):
    sc, so, sv, sn = slice_rows(chars, offsets, validity, 0, rows)
    col = nvstrings.from_offsets(sc, so, rows, sv, sn)
    res = torch.empty(rows, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        L.custr_contains_re(col.m_cptr, pat, res.data_ptr(), 1)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        L.custr_contains_re(col.m_cptr, pat, res.data_ptr(), 1)
    b.record()
    torch.cuda.synchronize()
    L.custr_set_profiling(1)
    ks = []
    for _ in range(10):
        L.custr_contains_re(col.m_cptr, pat, res.data_ptr(), 1)
        ks.append(float(L.custr_last_kernel_ms()))
    L.custr_set_profiling(0)
    print("rows %9d  chars %11d  call %.4f ms  kernel %.4f ms  tier %s" % (rows, int(so[-1] - so[0]), a.elapsed_time(b) / 20, float(np.median(ks)),
                                                                          L.custr_last_regex_tier().decode()), flush=True)
    del col, res

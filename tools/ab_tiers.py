#!/usr/bin/env python3
"""A/B of regex tiers on the C2 column in ONE process (same box, same clocks): median CUDA-event time of contains_re per tier.
    python tools/ab_tiers.py [--tiers 0,4] [--pattern P] [--reps 30]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from custrings_b200 import nvstrings  # noqa: E402
from custrings_b200._lib import lib  # noqa: E402
from custrings_b200.workloads import c2_corpus  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tiers", default="0,4")
ap.add_argument("--pattern", default=r"\b\w{4,}\b")
ap.add_argument("--reps", type=int, default=30)
ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--bytes", type=int, default=1 << 30)
ap.add_argument("--item-kib", type=int, default=0, help="0 = library default, -1 = default without the graded first round, 4..32 = fixed")
ap.add_argument("--jit", type=int, default=1, help="0 never, 1 default policy, 2 always")
a = ap.parse_args()
chars, offsets, validity, nulls = c2_corpus(a.rows, a.bytes)
col = nvstrings.from_offsets(chars, offsets, a.rows, validity, nulls)
res = torch.empty(a.rows, dtype=torch.uint8, device="cuda")
L = lib()
L.custr_set_profiling(1)
L.custr_set_item_kib(a.item_kib)
L.custr_set_jit(a.jit, 0)
tiers = [int(t) for t in a.tiers.split(",")]
times = {t: [] for t in tiers}
for rep in range(a.reps + 3):
    for t in tiers:
        L.custr_set_regex_tier(t)
        m = L.custr_contains_re(col.m_cptr, a.pattern.encode(), res.data_ptr(), 1)
        if rep >= 3:
            times[t].append(L.custr_last_kernel_ms())
L.custr_set_regex_tier(0)
for t in tiers:
    print("tier %d: median %.4f ms  min %.4f ms  (matches %d)" % (t, float(np.median(times[t])), min(times[t]), m))
print("jit launches", L.custr_jit_launch_count(), "| note:", L.custr_jit_note().decode())

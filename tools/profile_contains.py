#!/usr/bin/env python3
"""Minimal driver for ncu: builds the C2 column on cuda:0 and calls contains_re a few times.
    ncu ... python tools/profile_contains.py [--rows R --bytes B --calls K --pattern P --tier T]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from custrings_b200 import nvstrings  # noqa: E402
from custrings_b200._lib import lib  # noqa: E402
from custrings_b200.workloads import c2_corpus  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--bytes", type=int, default=1 << 30)
ap.add_argument("--calls", type=int, default=4)
ap.add_argument("--pattern", default=r"\b\w{4,}\b")
ap.add_argument("--tier", type=int, default=0)
ap.add_argument("--ascii", action="store_true", help="replace every non-ASCII byte by a letter")
ap.add_argument("--op", default="contains", choices=["contains", "count", "replace", "tokenize", "split_record"])
a = ap.parse_args()
chars, offsets, validity, nulls = c2_corpus(a.rows, a.bytes)
if a.ascii:
    chars[chars >= 0x80] = 101
col = nvstrings.from_offsets(chars, offsets, a.rows, validity, nulls)
res = torch.empty(a.rows, dtype=torch.uint8, device="cuda")
L = lib()
L.custr_set_regex_tier(a.tier)
res32 = torch.empty(a.rows, dtype=torch.int32, device="cuda")
for _ in range(a.calls):
    if a.op == "contains":
        m = L.custr_contains_re(col.m_cptr, a.pattern.encode(), res.data_ptr(), 1)
    elif a.op == "count":
        m = L.custr_count_re(col.m_cptr, a.pattern.encode(), res32.data_ptr(), 1)
    elif a.op == "replace":
        m = col.replace(a.pattern, "#").size()
    elif a.op == "tokenize":
        from custrings_b200 import nvtext
        m = nvtext.tokenize(col).size()
    else:
        m = col.split_record_flat(" ")[0].size()
torch.cuda.synchronize()
print("matches", m, "tier", L.custr_last_regex_tier().decode())

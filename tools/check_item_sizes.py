#!/usr/bin/env python3
"""contains_re / count_re on C2 prefixes with several forced work-item sizes against the exact Pike VM: a quick GPU check that
results do not depend on where the items start (the count-mode carry bug of round 2 showed up here first)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from custrings_b200 import nvstrings
from custrings_b200._lib import lib
from custrings_b200.workloads import c2_corpus
n, nbytes = 10_000_000, 1 << 30
chars, offsets, validity, nulls = c2_corpus(n, nbytes)
col = nvstrings.from_offsets(chars, offsets, n, validity, nulls)
L = lib()
PAT = rb"\b\w{4,}\b"
for rows in (1_000_000, 2_000_000):
    sub = col[0:rows]
    m = sub.size()
    def run(kib, tier):
        L.custr_set_item_kib(kib); L.custr_set_regex_tier(tier)
        hit = torch.zeros(m, dtype=torch.uint8, device="cuda"); cnt = torch.zeros(m, dtype=torch.int32, device="cuda")
        a = L.custr_contains_re(sub.m_cptr, PAT, hit.data_ptr(), 1)
        b = L.custr_count_re(sub.m_cptr, PAT, cnt.data_ptr(), 1)
        L.custr_set_regex_tier(0); L.custr_set_item_kib(0)
        return a, b, hit.cpu().numpy(), cnt.cpu().numpy(), L.custr_last_regex_tier().decode()
    ref = run(0, 1)
    print("rows", rows, "vm:", ref[0], ref[1])
    for kib in (32, 0, 16, 24, 20):
        r = run(kib, 0)
        dh = np.nonzero(r[2] != ref[2])[0]; dc = np.nonzero(r[3] != ref[3])[0]
        print(" kib", kib, "contains", r[0], "count", r[1], "hit diffs", len(dh), dh[:5], "count diffs", len(dc), dc[:5], r[3][dc[:5]], ref[3][dc[:5]], flush=True)
        if len(dc):
            o = offsets
            for i in dc[:3]:
                print("   row", i, "bytes", o[i], o[i+1], repr(bytes(chars[o[i]:o[i+1]]).decode(errors="replace"))[:160])

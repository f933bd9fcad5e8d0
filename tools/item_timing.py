#!/usr/bin/env python3
"""Per-warp time stamps of the item kernel (variant built with -DCUSTR_ITEM_TIMING, see tools/build_variant.py):
    python tools/build_variant.py T -DCUSTR_ITEM_TIMING
    CUSTR_LIB=custrings_b200/_variants/libcustr_T.so python tools/item_timing.py
For each column size: when the warps enter, how long phase 0 / A / B of their 1st, 2nd, ... item take, when they finish."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from custrings_b200 import nvstrings  # noqa: E402
from custrings_b200._lib import lib  # noqa: E402
from custrings_b200.workloads import c2_corpus, slice_rows  # noqa: E402

L = lib()
raw = C.CDLL(os.environ["CUSTR_LIB"])
SLOTS, WARPS = 4 + 4 * 6, 148 * 3 * 8
buf = np.zeros(WARPS * SLOTS, np.uint64)
n, nbytes = 10_000_000, 1 << 30
chars, offsets, validity, nulls = c2_corpus(n, nbytes)
pat = rb"\b\w{4,}\b"
for rows, kib in ((10_000, 32), (10_000, 0), (1_250_000, 32), (10_000_000, -1)):
    sc, so, sv, sn = slice_rows(chars, offsets, validity, 0, rows)
    col = nvstrings.from_offsets(sc, so, rows, sv, sn)
    res = torch.empty(rows, dtype=torch.uint8, device="cuda")
    L.custr_set_item_kib(kib)
    for _ in range(3):
        L.custr_contains_re(col.m_cptr, pat, res.data_ptr(), 1)
    torch.cuda.synchronize()
    raw.custr_dbg_item_times_clear()
    L.custr_contains_re(col.m_cptr, pat, res.data_ptr(), 1)
    got = raw.custr_dbg_item_times(buf.ctypes.data_as(C.c_void_p), buf.size)
    L.custr_set_item_kib(0)
    t = buf.reshape(WARPS, SLOTS).astype(np.int64)
    live = t[:, 0] > 0
    t0 = t[live, 0].min()
    us = lambda x: (x - t0) / 1e3  # noqa: E731
    q = lambda a: "p10 %.1f  med %.1f  p90 %.1f  max %.1f" % tuple(np.percentile(a, [10, 50, 90, 100]))  # noqa: E731
    print("== rows %d, item KiB %d: %d warps entered (%d words read)" % (rows, kib, int(live.sum()), got))
    print("   entry after first warp (us):      ", q(us(t[live, 0])))
    print("   first item known after entry (us):", q((t[live, 1] - t[live, 0]) / 1e3))
    for k in range(6):
        b = 4 + 4 * k
        has = live & (t[:, b + 3] > 0)
        if not has.any():
            break
        print("   item #%d (%d warps): start %s" % (k + 1, int(has.sum()), q(us(t[has, b]))))
        print("        phase 0 %s | phase A %s | phase B %s" % (q((t[has, b + 1] - t[has, b]) / 1e3), q((t[has, b + 2] - t[has, b + 1]) / 1e3),
                                                              q((t[has, b + 3] - t[has, b + 2]) / 1e3)))
        print("        end   %s" % q(us(t[has, b + 3])))
    if rows == 10_000_000:  # are the slow warps the warps of particular SMs?
        b = 4
        dur = (t[:, b + 3] - t[:, b]) / 1e3
        sm = t[:, 2]
        per_sm = {int(k): float(np.median(dur[(sm == k) & live])) for k in np.unique(sm[live])}
        vals = np.array(sorted(per_sm.values()))
        print("   item #1 duration, median per SM: min %.1f  p25 %.1f  med %.1f  p75 %.1f  max %.1f us over %d SMs" % (
            vals[0], np.percentile(vals, 25), np.median(vals), np.percentile(vals, 75), vals[-1], len(vals)))
        slow = sorted(per_sm, key=per_sm.get)[-12:]
        print("   slowest SMs:", [(k, round(per_sm[k], 1)) for k in slow])
        wid = np.arange(WARPS) % 8
        print("   item #1 duration by warp index in its CTA:", [round(float(np.median(dur[(wid == w) & live])), 1) for w in range(8)])
        cta_slot = {}
        print("   warps per SM:", np.bincount(sm[live].astype(np.int64)).tolist()[:20], "...")
    del col, res

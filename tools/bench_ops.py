#!/usr/bin/env python3
"""Secondary measurements (not the headline): device time of the other hot-path ops on BASELINE-shaped inputs.
Prints one JSON object per op: ms (median of reps, CUDA events), rows, chars, effective GB/s over chars read once."""
import json
import os
import sys
import random

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from custrings_b200 import nvstrings, nvcategory, nvtext  # noqa: E402
from custrings_b200._lib import lib  # noqa: E402
from custrings_b200.workloads import c2_corpus  # noqa: E402


def timed(fn, reps=5):
    fn()
    fn()  # two warm-up calls: the first one grows the stream-ordered memory pool
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
        del r
    return float(np.median(ts))


def report(name, ms, rows, chars, extra=None):
    d = {"op": name, "ms": round(ms, 3), "rows": rows, "chars": chars, "Mrows_per_s": round(rows / ms / 1e3, 1), "GBps_chars": round(chars / ms / 1e6, 1)}
    d.update(extra or {})
    print(json.dumps(d), flush=True)


def main():
    n = int(os.environ.get("ROWS", 10_000_000))
    nbytes = int(os.environ.get("BYTES", 1 << 30))
    chars, offsets, validity, nulls = c2_corpus(n, nbytes)
    col = nvstrings.from_offsets(chars, offsets, n, validity, nulls)
    res8 = torch.empty(n, dtype=torch.uint8, device="cuda")
    res32 = torch.empty(n, dtype=torch.int32, device="cuda")
    L = lib()
    pat = rb"\b\w{4,}\b"
    report("contains_re bitstream", timed(lambda: L.custr_contains_re(col.m_cptr, pat, res8.data_ptr(), 1)), n, nbytes)
    report("match ^\\w+ \\w+ bitstream", timed(lambda: L.custr_match(col.m_cptr, rb"\w+ \w+", res8.data_ptr(), 1)), n, nbytes)
    report("contains_re (a|b)c dag-bitstream", timed(lambda: L.custr_contains_re(col.m_cptr, rb"(ab|cd)e", res8.data_ptr(), 1)), n, nbytes)
    L.custr_set_regex_tier(1)
    report("contains_re pikevm", timed(lambda: L.custr_contains_re(col.m_cptr, pat, res8.data_ptr(), 1), 3), n, nbytes)
    L.custr_set_regex_tier(0)
    report("count_re (bit streams + word scans, last-loop chains)", timed(lambda: L.custr_count_re(col.m_cptr, pat, res32.data_ptr(), 1), 3), n, nbytes)
    report("replace_re \\b\\w{4,}\\b -> # (bit streams + word scans, 2 passes)", timed(lambda: col.replace(r"\b\w{4,}\b", "#"), 3), n, nbytes)
    report("replace literal 'ab' -> X (bit-stream splice)", timed(lambda: col.replace("ab", "X"), 3), n, nbytes)
    report("replace literal ' ' -> '_' (bit-stream splice)", timed(lambda: col.replace(" ", "_", regex=False), 3), n, nbytes)
    L.custr_set_regex_tier(2)
    try:
        report("replace literal 'ab' -> X (per-row walk)", timed(lambda: col.replace("ab", "X"), 3), n, nbytes)
    finally:
        L.custr_set_regex_tier(0)
    report("contains literal", timed(lambda: L.custr_contains(col.m_cptr, b"abcd", res8.data_ptr(), 1)), n, nbytes)
    report("find literal", timed(lambda: L.custr_find(col.m_cptr, b"abcd", 0, -1, res32.data_ptr(), 1)), n, nbytes)
    report("tokenize whitespace", timed(lambda: nvtext.tokenize(col), 3), n, nbytes)
    report("split_record ' ' (flat, row offsets copied to the host)", timed(lambda: col.split_record_flat(" "), 3), n, nbytes)
    import ctypes as C
    row_off_dev = torch.empty(n + 1, dtype=torch.int32, device="cuda")

    def split_dev(delim=b" "):
        tok = C.c_void_p()
        k = L.custr_split_record(col.m_cptr, delim, -1, C.byref(tok), row_off_dev.data_ptr(), 1)
        assert k >= 0, L.custr_last_error()
        L.custr_column_free(tok)
        return k
    l0 = L.custr_launch_count()
    ntok = split_dev()
    per_call = int(L.custr_launch_count() - l0)
    report("split_record ' ' (flat, device resident; bit streams)", timed(split_dev, 5), n, nbytes, {"tokens": int(ntok), "launches_per_call": per_call})
    L.custr_set_regex_tier(2)
    try:
        report("split_record ' ' (flat, device resident; per-row walk)", timed(split_dev, 3), n, nbytes)
    finally:
        L.custr_set_regex_tier(0)
    sub = col[0:1_000_000]
    report("split(' ', n=7) 8 columns, 1M rows", timed(lambda: sub.split(" ", 7), 3), 1_000_000, int(sub.byte_count()))
    report("hash", timed(lambda: L.custr_hash(col.m_cptr, res32.data_ptr(), 1)), n, nbytes)
    # C3: README chain on a 10M-row day-of-week column
    rng = np.random.Generator(np.random.PCG64(7))
    days = [b"Sun", b"Mon", b"Tues", b"Wed", b"Thur", b"Fri", b"Sat"]
    pick = rng.integers(0, 7, size=n)
    dl = np.array([len(d) for d in days])
    off = np.zeros(n + 1, np.int64)
    np.cumsum(dl[pick], out=off[1:])
    dchars = np.frombuffer(b"".join(days), np.uint8)
    doff = np.zeros(8, np.int64)
    np.cumsum(dl, out=doff[1:])
    idx = np.repeat(doff[pick] - off[:-1], dl[pick]) + np.arange(off[-1])
    dcol = nvstrings.from_offsets(dchars[idx], off.astype(np.int32), n)

    def chain():
        c = dcol
        for i, d in enumerate(days):
            c = c.replace(d.decode(), str(i))
        return c
    report("C3 README 7-step replace chain (regex=True -> literal kernel)", timed(chain, 3), n, int(off[-1]))
    # C4: category, 1000 keys
    k = 1000
    alphabet = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz0123456789", np.uint8)
    keys = list(dict.fromkeys(alphabet[rng.integers(0, 36, size=int(l))].tobytes() for l in rng.integers(8, 25, size=k)))
    kl = np.array([len(x) for x in keys])
    koff = np.zeros(len(keys) + 1, np.int64)
    np.cumsum(kl, out=koff[1:])
    kchars = np.frombuffer(b"".join(keys), np.uint8)
    pick = rng.integers(0, len(keys), size=n)
    off = np.zeros(n + 1, np.int64)
    np.cumsum(kl[pick], out=off[1:])
    idx = np.repeat(koff[pick] - off[:-1], kl[pick]) + np.arange(off[-1])
    ccol = nvstrings.from_offsets(kchars[idx], off.astype(np.int32), n)
    report("C4 nvcategory.from_strings (1000 keys)", timed(lambda: nvcategory.from_strings(ccol), 3), n, int(off[-1]))


if __name__ == "__main__":
    main()

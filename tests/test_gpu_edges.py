"""Edge cases of the column layout against the bitstream kernels (windowing, row bookkeeping, views): every result
is compared with the exact Pike-VM tier and, where cheap, with the oracle."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu
PATS = [r"\b\w{4,}\b", r"\d+", r"^a", r"z$", r"é+x", r"\bq", r"[a-c]{2}\s", r"ab|cd", r"x\b", r"^$"]


def _both(col, pat, anchored=False):
    from custrings_b200._lib import lib
    out = {}
    for tier in (0, 1, 2, 3):
        lib().custr_set_regex_tier(tier)
        try:
            out[tier] = col.match(pat) if anchored else col.contains(pat)
        finally:
            lib().custr_set_regex_tier(0)
    return out


def _check(strs, oracle=None):
    from custrings_b200 import nvstrings
    col = nvstrings.to_device(strs)
    for pat in PATS:
        for anchored in (False, True):
            r = _both(col, pat, anchored)
            assert r[0] == r[1] == r[2] == r[3], (pat, anchored)
            if oracle is not None:
                ref = oracle.RefStrings.from_list(strs)
                want = (ref.match(pat) if anchored else ref.contains_re(pat))[0].tolist()
                assert [False if x is None else x for x in r[0]] == want, (pat, anchored)
    return col


def test_degenerate_columns(oracle):
    _check([""], oracle)
    _check([None], oracle)
    _check(["", "", None, ""] * 100, oracle)
    _check(["abcd"], oracle)
    _check(["a"] * 5000, oracle)                       # > 32 rows per window: multi-chunk row bookkeeping
    _check(["", "abcd", ""] * 3000, oracle)
    _check([None, "word", "", "zz z", "é"] * 2000, oracle)


def test_long_rows_cross_many_windows(oracle):
    big = "ab " * 300_000 + "abcd"                      # ~900 KB row, the only match at the very end
    rows = ["x", big, "", "q" * 5000 + " éx", "12 " * 20000, None, "tail abcd"]
    col = _check(rows, oracle)
    assert col.contains(r"\b\w{4,}\b") == [False, True, False, True, False, None, True]


def test_window_boundary_alignment(oracle):
    # rows whose ends / multi-byte characters land exactly on 1024 / 2048 byte window boundaries
    rows = []
    for total in (1023, 1024, 1025, 2047, 2048, 2049, 4096):
        rows.append("a" * (total - 5) + " bcde")
    rows.append("a" * 2047 + "é" + "bcd")              # 2-byte char straddling a 2048-byte boundary
    rows.append("x" * 1022 + "日本語" + "y" * 10)        # 3-byte chars across a 1024-byte boundary
    rows.append("w" * 4095)
    rows += ["ab", "", None, "é" * 700, "abc é" * 500]
    _check(rows, oracle)


def test_row_slice_views_and_gather():
    from custrings_b200 import nvstrings
    import random
    rng = random.Random(5)
    strs = ["".join(rng.choice("ab cdé1_\n") for _ in range(rng.choice([0, 1, 3, 9, 40, 200]))) for _ in range(5000)]
    strs[10] = None
    col = nvstrings.to_device(strs)
    for lo, hi in ((0, 5000), (1, 4999), (7, 3001), (2500, 2501), (4999, 5000), (100, 100)):
        view = col[lo:hi]
        for pat in PATS[:6]:
            r = _both(view, pat)
            assert r[0] == r[1] == r[3], (lo, hi, pat)
            assert r[0] == _both(nvstrings.to_device(strs[lo:hi]), pat)[1] if hi > lo else r[0] == []


def test_adopted_device_buffers():
    import torch
    from custrings_b200 import nvstrings
    strs = ["alpha beta", "", "xy", "gamma1234 z", "é", "a_b1 c"] * 1000
    enc = [s.encode() for s in strs]
    offsets = np.zeros(len(enc) + 1, np.int32)
    np.cumsum([len(e) for e in enc], out=offsets[1:])
    chars = torch.from_numpy(np.frombuffer(b"".join(enc), np.uint8).copy()).cuda()
    offs = torch.from_numpy(offsets).cuda()
    col = nvstrings.from_device_view(chars, offs, len(strs))
    want = nvstrings.to_device(strs).contains(r"\b\w{4,}\b")
    assert col.contains(r"\b\w{4,}\b") == want
    assert col.to_host() == strs


def test_ipc_export_import_between_processes(tmp_path):
    """CUDA IPC: a column exported in this process is opened, zero copy, by a second process on the same GPU
    (reference ipc_transfer.h / create_ipc_transfer / create_from_ipc)"""
    import subprocess
    import sys
    from custrings_b200 import nvstrings
    rows = ["héllo", None, "", "wörld 123", "x" * 300] * 50
    col = nvstrings.to_device(rows)
    data = col.get_ipc_data()
    (tmp_path / "h.bin").write_bytes(data)
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from custrings_b200 import nvstrings\n"
            "c = nvstrings.create_from_ipc(open(%r, 'rb').read())\n"
            "import json; json.dump({'rows': c.to_host(), 'hits': c.contains('\\\\d+')}, open(%r, 'w'))\n"
            % (ROOT, str(tmp_path / "h.bin"), str(tmp_path / "out.json")))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    out = json.load(open(tmp_path / "out.json"))
    assert out["rows"] == rows
    assert out["hits"] == col.contains(r"\d+")


def test_stream_paths_on_row_slice_views(oracle):
    """The bit-stream forms of replace / replace_re / split_record / tokenize / find on row-slice VIEWS (first offset not zero and
    not aligned, the bytes of the parent's neighbouring rows right next to the view's) against the oracle on the same rows.
    The last row of the view ends where the parent's next row — which would complete an occurrence — begins."""
    from custrings_b200 import nvstrings, nvtext
    import random
    rng = random.Random(9)
    words = ["alpha", "be", "gamma7", "a", "ab", "b", "abab", "é", "x_y", "zz ", "delta,", ""]
    strs = [" ".join(rng.choice(words) for _ in range(rng.choice([1, 3, 8, 20, 60]))) for _ in range(6000)]
    strs[11] = None
    for k in range(50, 6000, 97):  # "...a" | "b..." : 'ab' straddles the row boundary, "wor" | "ds": a word run does too
        strs[k] = strs[k] + " xa"
        strs[k + 1] = "b wor" if k % 2 else "bcde" + strs[k + 1]
    col = nvstrings.to_device(strs)
    for lo, hi in ((0, 6000), (51, 148), (7, 5001), (148, 149), (1000, 1000 + 2911)):
        view = col[lo:hi]
        ref = oracle.RefStrings.from_list(strs[lo:hi])
        dec = lambda r: [None if x is None else x.decode() for x in r.to_list()]  # noqa: E731
        assert view.replace("ab", "<>", regex=False).to_host() == dec(ref.replace("ab", "<>")), (lo, hi)
        assert view.replace(" ", "", regex=False).to_host() == dec(ref.replace(" ", "")), (lo, hi)
        assert view.replace(r"\b\w{4,}\b", "#").to_host() == dec(ref.replace_re(r"\b\w{4,}\b", "#")), (lo, hi)
        assert view.replace(r"[a-z]+", "").to_host() == dec(ref.replace_re(r"[a-z]+", "")), (lo, hi)
        got = view.split_record(" ")
        want = ref.split_record(" ")[0]
        assert [None if g is None else g.to_host() for g in got] == [None if w is None else dec(w) for w in want], (lo, hi)
        assert nvtext.tokenize(view).to_host() == dec(ref.tokenize()), (lo, hi)
        f = view.find("ab")
        assert [g for g in f if g is not None] == [int(x) for x, g in zip(ref.find("ab")[0], f) if g is not None], (lo, hi)


def test_item_starts_inside_windows(oracle):
    """Work items whose first row starts deep inside a 2 KiB window (odd item sizes, row-slice views, small columns whose item size
    is derived from the column size): count_re once carried the match bits of the previous item's rows into a first row that
    did not end in its first window.  Every bit-stream path against the oracle with item sizes of 5 / 7 / 13 KiB and the default."""
    from custrings_b200 import nvstrings, nvtext
    from custrings_b200._lib import lib
    import random
    rng = random.Random(21)
    words = ["alpha", "be", "gamma7", "a", "ab", "b", "x_y", "é", "zz", "delta,", "12345", ""]
    strs = []
    for k in range(4000):
        nw = rng.choice([1, 3, 8, 20, 60, 400, 900])  # rows of up to ~5 KB: many do not end in the window they start in
        strs.append(None if k % 211 == 3 else " ".join(rng.choice(words) for _ in range(nw)))
    col = nvstrings.to_device(strs)
    ref = oracle.RefStrings.from_list(strs)
    dec = lambda r: [None if x is None else x.decode() for x in r.to_list()]  # noqa: E731
    want = {
        "count": [int(x) for x in ref.count_re(r"\b\w{4,}\b")[0]],
        "count2": [int(x) for x in ref.count_re(r"\d+")[0]],
        "contains": [bool(x) for x in ref.contains_re(r"\b\w{4,}\b")[0]],
        "replace_re": dec(ref.replace_re(r"\b\w{4,}\b", "#")),
        "replace": dec(ref.replace("ab", "<>")),
        "tokenize": dec(ref.tokenize()),
    }
    valid = [s is not None for s in strs]
    for kib in (0, 5, 7, 13, 32):
        lib().custr_set_item_kib(kib)
        try:
            for view_lo in (0, 17):
                v = col[view_lo:len(strs)] if view_lo else col
                ok = valid[view_lo:]
                assert [c for c, k in zip(v.count(r"\b\w{4,}\b"), ok) if k] == [c for c, k in zip(want["count"][view_lo:], ok) if k], (kib, view_lo)
                assert [c for c, k in zip(v.count(r"\d+"), ok) if k] == [c for c, k in zip(want["count2"][view_lo:], ok) if k], (kib, view_lo)
                assert [c for c, k in zip(v.contains(r"\b\w{4,}\b"), ok) if k] == [c for c, k in zip(want["contains"][view_lo:], ok) if k], (kib, view_lo)
                assert v.replace(r"\b\w{4,}\b", "#").to_host() == want["replace_re"][view_lo:], (kib, view_lo)
                assert v.replace("ab", "<>", regex=False).to_host() == want["replace"][view_lo:], (kib, view_lo)
            assert nvtext.tokenize(col).to_host() == want["tokenize"], kib
        finally:
            lib().custr_set_item_kib(0)


def test_release_cached_memory_between_calls():
    """custr_release_cached_memory hands the big-block cache and the pool's unused memory back; calls after it allocate afresh."""
    from custrings_b200 import nvstrings
    from custrings_b200._lib import lib
    rows = ["alpha beta gamma delta " * 40] * 20000  # ~18 MB in, ~18 MB out: above and below the 32 MiB cache threshold mix
    col = nvstrings.to_device(rows + rows)
    first = col.replace("alpha", "A", regex=False).to_host()
    lib().custr_release_cached_memory()
    again = col.replace("alpha", "A", regex=False).to_host()
    assert first == again and again[0].startswith("A beta")
    lib().custr_release_cached_memory()
    assert col.contains(r"\bgamma\b")[:2] == [True, True]

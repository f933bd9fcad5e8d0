"""Full-size checks (BASELINE.json configs) through size-independent properties, where the oracle would take minutes:
tier agreement, count/contains consistency, replace idempotence, tokenize/split agreement, dictionary round trip."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
PAT = r"\b\w{4,}\b"


@pytest.fixture(scope="module")
def c2():
    import torch
    from custrings_b200 import nvstrings
    from custrings_b200.workloads import c2_corpus
    n, nbytes = 10_000_000, 1 << 30
    chars, offsets, validity, nulls = c2_corpus(n, nbytes)
    col = nvstrings.from_offsets(chars, offsets, n, validity, nulls)
    return n, chars, offsets, validity, nulls, col


def _bool(col, pat, tier, anchored=False):
    import torch
    from custrings_b200._lib import lib
    out = torch.zeros(col.size(), dtype=torch.uint8, device="cuda")
    lib().custr_set_regex_tier(tier)
    try:
        fn = lib().custr_match if anchored else lib().custr_contains_re
        cnt = fn(col.m_cptr, pat.encode(), out.data_ptr(), 1)
        used = lib().custr_last_regex_tier().decode()
    finally:
        lib().custr_set_regex_tier(0)
    return out, cnt, used


def test_c2_tiers_agree_full_size(c2, oracle):
    import torch
    n, chars, offsets, validity, nulls, col = c2
    fast, cnt_fast, used = _bool(col, PAT, 0)
    assert used == "bitstream"
    exact, cnt_exact, used1 = _bool(col, PAT, 1)
    assert used1 == "pikevm"
    assert cnt_fast == cnt_exact == int(fast.sum().item())
    assert torch.equal(fast, exact)
    for tier in (2, 3):
        other, cnt, _ = _bool(col, PAT, tier)
        assert cnt == cnt_exact and torch.equal(other, exact)
    # oracle on a prefix of the same column
    from custrings_b200.workloads import slice_rows
    m = 200_000
    c, o, v, nn = slice_rows(chars, offsets, validity, 0, m)
    want, wcnt = oracle.RefStrings.from_arrays(c, o, v, nn).contains_re(PAT)
    assert np.array_equal(fast[:m].cpu().numpy().astype(bool), want)
    # null rows are false
    valid = np.unpackbits(validity, bitorder="little")[:n].astype(bool)
    assert not fast.cpu().numpy()[~valid].any()


def test_c2_more_patterns_tiers_agree(c2):
    import torch
    n, chars, offsets, validity, nulls, col = c2
    sub = col[0:2_000_000]
    for pat, anchored in ((r"\d+", False), (r"[a-f]{3}\b", False), (r"é", False), (r"\w+ \w+", True), (r"^\w{8}", False), (r"z\w*$", False),
                          (r"\bq\w+|\bz\w+", False)):
        a, ca, _ = _bool(sub, pat, 0, anchored)
        b, cb, _ = _bool(sub, pat, 1, anchored)
        assert ca == cb and torch.equal(a, b), pat


def test_c2_count_and_replace_properties(c2):
    import torch
    from custrings_b200._lib import lib
    n, chars, offsets, validity, nulls, col = c2
    sub = col[0:1_000_000]
    m = sub.size()
    hit, cnt, _ = _bool(sub, PAT, 0)
    counts = torch.zeros(m, dtype=torch.int32, device="cuda")
    nz = lib().custr_count_re(sub.m_cptr, PAT.encode(), counts.data_ptr(), 1)
    assert nz == cnt and torch.equal(counts > 0, hit.bool())
    replaced = sub.replace(PAT, "#")
    again, cnt2, _ = _bool(replaced, PAT, 0)
    assert cnt2 == 0 and not again.any()
    # byte accounting: every match shrinks the row by (len(match) - 1) >= 3 bytes
    assert replaced.byte_count() <= sub.byte_count() - 3 * int(counts.sum().item())
    assert replaced.size() == m and replaced.null_count() == sub.null_count()


def test_c2_tokenize_split_agree(c2):
    from custrings_b200 import nvtext
    import torch
    n, chars, offsets, validity, nulls, col = c2
    sub = col[0:1_000_000]
    tc = torch.zeros(sub.size(), dtype=torch.int32, device="cuda")
    nvtext.token_count(sub, None, devptr=tc.data_ptr())
    tokens = nvtext.tokenize(sub)
    assert tokens.size() == int(tc.sum().item())
    flat, row_off = sub.split_record_flat(None)
    # whitespace split_record: one "" token for valid rows without tokens, else the same tokens as tokenize
    valid = np.unpackbits(sub.to_arrays()[2], bitorder="little")[: sub.size()].astype(bool)
    per_row = np.diff(row_off)
    tcn = tc.cpu().numpy()
    assert np.array_equal(per_row[valid], np.maximum(tcn[valid], 1)) and (per_row[~valid] == 0).all()
    assert tokens.byte_count() == flat.byte_count()


def test_c4_category_roundtrip_large():
    import torch
    from custrings_b200 import nvstrings, nvcategory
    rng = np.random.Generator(np.random.PCG64(11))
    k, n = 1000, 5_000_000
    lens = rng.integers(8, 25, size=k)
    alphabet = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz0123456789", np.uint8)
    keys = [alphabet[rng.integers(0, 36, size=l)].tobytes() for l in lens]
    keys = list(dict.fromkeys(keys))
    pick = rng.integers(0, len(keys), size=n)
    klen = np.array([len(x) for x in keys])
    koff = np.zeros(len(keys) + 1, np.int64)
    np.cumsum(klen, out=koff[1:])
    kchars = np.frombuffer(b"".join(keys), np.uint8)
    row_len = klen[pick]
    offsets = np.zeros(n + 1, np.int64)
    np.cumsum(row_len, out=offsets[1:])
    idx = np.repeat(koff[pick] - offsets[:-1], row_len) + np.arange(offsets[-1])
    chars = kchars[idx]
    col = nvstrings.from_offsets(chars, offsets.astype(np.int32), n)
    cat = nvcategory.from_strings(col)
    got_keys = [x.encode() for x in cat.keys().to_host()]
    assert got_keys == sorted(keys)
    rank = {key: i for i, key in enumerate(got_keys)}
    want = np.array([rank[x] for x in keys], np.int32)[pick]
    vals = torch.zeros(n, dtype=torch.int32, device="cuda")
    cat.values(devptr=vals.data_ptr())
    assert np.array_equal(vals.cpu().numpy(), want)


def test_c1_split_csv_shape(oracle):
    from custrings_b200 import nvstrings
    rng = np.random.Generator(np.random.PCG64(3))
    rows = ["%d,%s,%d.%02d,%s,,x,%d,Sacramento,CA,%05d,Residential,%d" %
            (i, "ab" * int(rng.integers(0, 4)), rng.integers(0, 99), rng.integers(0, 99), "é" if i % 7 == 0 else "q", rng.integers(1, 6),
             rng.integers(0, 99999), rng.integers(1000, 900000)) for i in range(985)]  # 985 rows x 12 columns like data/985-rows.csv
    cols = nvstrings.to_device(rows).split(",")
    want = oracle.RefStrings.from_list(rows).split(",")
    assert len(cols) == len(want) == 12
    for a, b in zip(cols, want):
        assert oracle.unpack(*a.to_arrays()) == b.to_list()


def test_split_record_bitstream_equals_per_row_path(c2, oracle):
    """split_record with one delimiter byte through the bit-stream compaction (split_bits.cuh) against this repo's per-row path
    on 2 M rows of C2 (many work items, rows straddling windows, null rows) and against the oracle on a prefix; a column
    with empty valid rows must fall back and still be right."""
    from custrings_b200 import nvstrings
    from custrings_b200._lib import lib
    from custrings_b200.workloads import slice_rows
    n, chars, offsets, validity, nulls, col = c2
    sub = col[0:2_000_000]
    for delim in (" ", "e"):
        tok_b, ro_b = sub.split_record_flat(delim)
        lib().custr_set_regex_tier(2)  # bit-stream paths off: per-row kernels
        try:
            tok_r, ro_r = sub.split_record_flat(delim)
        finally:
            lib().custr_set_regex_tier(0)
        assert np.array_equal(ro_b, ro_r), delim
        for a, b in zip(tok_b.to_arrays(), tok_r.to_arrays()):
            assert np.array_equal(a, b), delim
    m = 50_000
    c, o, v, nn = slice_rows(chars, offsets, validity, 0, m)
    want, total = oracle.RefStrings.from_arrays(c, o, v, nn).split_record(" ")
    got = nvstrings.from_offsets(c, o, m, v, nn).split_record(" ")
    assert [None if g is None else g.to_host() for g in got] == [None if w is None else [x.decode() for x in w.to_list()] for w in want]
    rows = ["a b", "", None, " x ", "", "tail"] * 4000  # empty VALID rows: not expressible as bits, must take the per-row path
    got = nvstrings.to_device(rows).split_record(" ")
    want = oracle.RefStrings.from_list(rows).split_record(" ")[0]
    assert [None if g is None else g.to_host() for g in got] == [None if w is None else [x.decode() for x in w.to_list()] for w in want]


def test_literal_replace_bitstream_equals_per_row_path(c2, oracle):
    """replace with a literal, border-free target through the bit-stream splice (replace_bits.cuh) against this repo's per-row
    path on 2 M rows of C2 and against the oracle on a prefix; then targeted layouts: occurrences straddling lane / window /
    row boundaries, rows longer than many windows, empty and null rows, multi-byte targets, growing and shrinking replacements,
    and a bordered target ("aa"), which must take the per-row path and still be right."""
    from custrings_b200 import nvstrings
    from custrings_b200._lib import lib
    from custrings_b200.workloads import slice_rows
    n, chars, offsets, validity, nulls, col = c2

    def both(column, tgt, repl):
        got = column.replace(tgt, repl, regex=False)
        lib().custr_set_regex_tier(2)  # bit-stream paths off: per-row kernel
        try:
            ref = column.replace(tgt, repl, regex=False)
        finally:
            lib().custr_set_regex_tier(0)
        for a, b in zip(got.to_arrays(), ref.to_arrays()):
            assert np.array_equal(a, b), (tgt, repl)
        return got

    sub = col[0:2_000_000]
    for tgt, repl in ((" ", "_"), ("e", ""), ("th", "THE"), ("ing ", "#"), ("a", "xyz")):
        both(sub, tgt, repl)
    m = 50_000
    c, o, v, nn = slice_rows(chars, offsets, validity, 0, m)
    want = oracle.RefStrings.from_arrays(c, o, v, nn).replace("he", "<HE>")
    got = nvstrings.from_offsets(c, o, m, v, nn).replace("he", "<HE>", regex=False)
    assert got.to_host() == [None if x is None else x.decode() for x in want.to_list()]

    rng = np.random.default_rng(11)
    rows = []
    for k in range(3000):
        ln = int(rng.choice([0, 1, 2, 3, 30, 63, 64, 65, 127, 1983, 1984, 1985, 2047, 2048, 2049, 5000]))
        rows.append(None if k % 97 == 5 else "".join(rng.choice(list("abcé ,"), ln).tolist()))
    rows.append("ab" * 40_000)                      # one row over many windows, every byte replaced
    rows.append("x" * 100_000 + "abc")
    rows += ["abc", "", "bc", "ab", "cab", None, "abcabc"] * 50
    dev = nvstrings.to_device(rows)
    ref = oracle.RefStrings.from_list(rows)
    for tgt, repl in (("ab", "X"), ("abc", ""), ("c", "cc"), ("é", "e"), ("b,", "béb"), (", ", ""), ("ca", "123456"), ("é ", "")):
        got = both(dev, tgt, repl)
        assert got.to_host() == [None if x is None else x.decode() for x in ref.replace(tgt, repl).to_list()], (tgt, repl)
    for tgt, repl in (("aa", "b"), ("abab", "-")):  # bordered targets: per-row path
        got = dev.replace(tgt, repl, regex=False)
        assert got.to_host() == [None if x is None else x.decode() for x in ref.replace(tgt, repl).to_list()], (tgt, repl)


def test_replace_re_bitsplice_equals_walk_and_vm(c2, oracle):
    """replace_re of single-class chains (x{n,} / x+ with assertions) through the streaming splice over the chain kernel's span
    streams (replace_bits.cuh MODE 1) against the per-row walk of the same streams (tier 3), the exact Pike VM (tier 1) and
    the oracle: C2 rows, then rows built to put runs across lane / window / row boundaries, '_' inside \\w runs (a \\b that
    holds in the middle of a run), multi-byte characters in front of and inside matches, empty and null rows."""
    from custrings_b200 import nvstrings
    from custrings_b200._lib import lib
    from custrings_b200.workloads import slice_rows
    n, chars, offsets, validity, nulls, col = c2

    def tiers(column, pat, repl, which=(0, 3)):
        res = []
        for t in which:
            lib().custr_set_regex_tier(t)
            try:
                res.append(column.replace(pat, repl))
            finally:
                lib().custr_set_regex_tier(0)
        for r in res[1:]:
            for a, b in zip(res[0].to_arrays(), r.to_arrays()):
                assert np.array_equal(a, b), (pat, repl)
        return res[0]

    sub = col[0:1_000_000]
    for pat, repl in ((r"\b\w{4,}\b", "#"), (r"\d+", ""), (r"\s+", " "), (r"[a-m]{2,}", "<>"), (r"\w+\b", "word!")):
        tiers(sub, pat, repl)
    m = 30_000
    c, o, v, nn = slice_rows(chars, offsets, validity, 0, m)
    want = oracle.RefStrings.from_arrays(c, o, v, nn).replace_re(r"\b\w{4,}\b", "#")
    got = nvstrings.from_offsets(c, o, m, v, nn).replace(r"\b\w{4,}\b", "#")
    assert got.to_host() == [None if x is None else x.decode() for x in want.to_list()]

    rng = np.random.default_rng(12)
    rows = []
    for k in range(2500):
        ln = int(rng.choice([0, 1, 2, 3, 4, 5, 30, 63, 64, 65, 127, 1983, 1984, 1985, 2047, 2048, 2049, 4000]))
        rows.append(None if k % 89 == 7 else "".join(rng.choice(list("abc1_é日 ,"), ln).tolist()))
    rows.append("word " * 30_000)
    rows.append("w" * 70_000 + " tail")              # one run over many windows
    rows.append("abcd_" * 20_000)                    # \b inside every run
    rows += ["abcd", "", "abc", "_abcd_", None, "éabcd", "abcdé x", "日本語日本語"] * 40
    dev = nvstrings.to_device(rows)
    ref = oracle.RefStrings.from_list(rows)
    for pat, repl in ((r"\b\w{4,}\b", "#"), (r"\w{3,}", ""), (r"[a-c]+", "ABC"), (r"\b[a-c1]{2,}", "_"), (r"\w+\b", "é"), (r"\s+", "")):
        got = tiers(dev, pat, repl, which=(0, 3, 1))
        assert got.to_host() == [None if x is None else x.decode() for x in ref.replace_re(pat, repl).to_list()], (pat, repl)


def test_find_with_contains_prefilter_equals_plain_scan(c2, oracle):
    """find / rfind on a large column first ask the chain kernel which rows hold the needle at all; the positions must equal the
    plain per-row scan (tier 2 switches the pre-filter off) for any start / end, and the oracle's on a prefix."""
    from custrings_b200._lib import lib
    n, chars, offsets, validity, nulls, col = c2
    sub = col[0:2_000_000]
    for needle, start, end in (("abcd", 0, None), ("the", 0, None), ("e", 3, 40), ("é", 0, None), (" a", 2, None), ("zzzzq", 0, None)):
        for right in (False, True):
            fn = sub.rfind if right else sub.find
            got = fn(needle, start, end)
            lib().custr_set_regex_tier(2)
            try:
                ref = fn(needle, start, end)
            finally:
                lib().custr_set_regex_tier(0)
            assert got == ref, (needle, start, end, right)
    m = 50_000
    from custrings_b200.workloads import slice_rows
    want = oracle.RefStrings.from_arrays(*slice_rows(chars, offsets, validity, 0, m)).find("the")[0]
    got = col[0:m].find("the")
    assert [g for g in got if g is not None] == [int(x) for x, g in zip(want, got) if g is not None]

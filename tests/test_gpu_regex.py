"""GPU parity: regex entry points through the C-ABI (ctypes -> libcustr.so) vs the reference oracle."""
import random

import numpy as np
import pytest

from tests import corpus

pytestmark = pytest.mark.gpu

SPAN_PATTERNS = [r"\b\w{4,}\b", r"\d+", r"[a-z]+", r"ab", r"a\d+", r"é+", r"\w+$", r"^\w+", r"x\b", r"\w\w\w", r"日\w+", r"\bé\w*", r"[^a-c]+",
                 r"\W+", r"\s\S+", r"\B\w+", r"l+", r"\w{2,}\b", r"\Aa\w*", r"\d\d?", r"[é-ü]+", r"😀.", r".\b", r"\w+\Z", r"\D{3}", r"_+",
                 # every shape-specialised chain kernel (regex_chain64.cuh SPEC 1..4, NS 1..4)
                 r"\w+", r"\w{2,}", r"\w{3,}", r"\w{4,}", r"\d{2,}", r"\d{3,}", r"\d\d\d\d+", r"\b\w+\b", r"\b\w{2,}\b", r"\b\w{3,}\b",
                 r"\b\d+\b", r"\b\d{2,}\b", r"\b\d{3,}\b", r"\b\d{4,}\b", r"\s+", r"\s{2,}", r"[a-z]+", r"[a-z]{2,}", r"[a-z]{3,}", r"[a-z]{4,}"]


# top-level alternations of chains: OR of chain-kernel runs instead of the DAG interpreter
# chains with optional steps / early exits and with 5..8 classes
WIDE_CHAINS = [r"colou?r", r"https?://", r"warning", r"abcdefgh", r"qu?ick\s+brown", r"\d{1,3}", r"\d{2,4}:\d\d?", r"ab*c", r"a?b?c", r"x\d{2,4}\b",
               r"\bjump(s|ed)?", r"la+zy dog", r"[Tt]he quick"]
ALTERNATIONS = [r"\bthe\b|\bfox\b", r"(\bin\b)|(\ba\b)|(\bthe\b)", r"\d+|é", r"^a|b$", r"ab|cd|ef|gh", r"\w+@|\s\d", r"[a-c]x|[|]y|\|z", r"日|😀|é+"]


@pytest.fixture(scope="module")
def cols(oracle):
    from custrings_b200 import nvstrings
    rng = random.Random(7)
    strs = corpus.STRINGS + corpus.random_strings(rng, 400)
    return strs, nvstrings.to_device(strs), oracle.RefStrings.from_list(strs)


def _none_to(v, lst):
    return [v if x is None else x for x in lst]


@pytest.mark.parametrize("tier", [0, 1, 2, 3, 4])
def test_contains_match_count_patterns(cols, tier):
    from custrings_b200._lib import lib
    strs, dev, ref = cols
    lib().custr_set_regex_tier(tier)
    try:
        pats = [p for p in corpus.PATTERNS if p not in (r"(a|b)*c", r"((a|b)c)*d", "a+*")] + corpus.random_patterns(11, 150) + SPAN_PATTERNS[-20:] + ALTERNATIONS + WIDE_CHAINS
        for p in pats:
            rc, rn = ref.contains_re(p)
            assert _none_to(False, dev.contains(p)) == rc.tolist(), p
            rm, _ = ref.match(p)
            assert _none_to(False, dev.match(p)) == rm.tolist(), p
            rk, _ = ref.count_re(p)
            assert _none_to(0, dev.count(p)) == rk.tolist(), p
    finally:
        lib().custr_set_regex_tier(0)


def test_replace_re(cols):
    strs, dev, ref = cols
    pats = [r"\b\w{4,}\b", r"\d+", "a*", "l+", r"\s", "é", r"[^a-c]+", r"\bthe\b|\bfox\b", r"\w+", "x*?y", "^", "$", r"\b"] + corpus.random_patterns(5, 60)
    for p in pats:
        for repl, mx in (("<>", -1), ("", 2), ("é日", 1)):
            want = ref.replace_re(p, repl, mx).to_list()
            got = dev.replace(p, repl, mx).to_arrays()
            from oracle.ref import unpack
            assert unpack(*got) == want, (p, repl, mx)


def test_replace_re_multi(cols, oracle):
    from custrings_b200 import nvstrings
    strs, dev, ref = cols
    pats = [r"\d+", "[tT]he", r"\s+", "é"]
    repls = ["#", "THE", "_", "e"]
    want = ref.replace_re_multi(pats, oracle.RefStrings.from_list(repls)).to_list()
    got = dev.replace_multi(pats, nvstrings.to_device(repls), regex=True)
    assert oracle.unpack(*got.to_arrays()) == want
    want1 = ref.replace_re_multi(pats, oracle.RefStrings.from_list(["."])).to_list()
    got1 = dev.replace_multi(pats, ".", regex=True)
    assert oracle.unpack(*got1.to_arrays()) == want1


def test_roundtrip_and_attrs(cols, oracle):
    strs, dev, ref = cols
    assert dev.size() == len(strs)
    a = dev.to_arrays()
    r = ref.to_arrays()
    for x, y in zip(a, r):
        assert np.array_equal(x, y)
    assert dev.to_host() == strs
    assert _none_to(-1, dev.len()) == ref.len()[0].tolist()
    assert _none_to(0, dev.hash()) == ref.hash()[0].tolist()
    assert dev.null_count() == sum(s is None for s in strs)
    sub = dev[3:20]
    assert sub.to_host() == strs[3:20]
    assert dev.gather([5, 0, 14, 2]).to_host() == [strs[5], strs[0], strs[14], strs[2]]


def test_findall_extract_capture_spans(cols, oracle):
    """'next' rows of the scope table: findall / findall_record / extract / extract_record vs the oracle"""
    strs, dev, ref = cols

    def lst(c):
        return oracle.unpack(*c.to_arrays())

    for p in [r"\w\d", r"\d+", r"[a-z]+", "é", r"\b\w{4,}\b", r"a*", r"q", r"(\w+) (\w+)", r"l+|o"] + corpus.random_patterns(31, 25):
        want = [c.to_list() for c in ref.findall(p)]
        got = [lst(c) for c in dev.findall(p)]
        assert got == want, ("findall", p)
        wr = [r.to_list() for r in ref.findall_record(p)]
        gr = [lst(r) for r in dev.findall_record(p)]
        assert gr == wr, ("findall_record", p)
    for p in [r"(\w+) (\w+)", r"(\w)(\d)?", r"(x*)(y)", r"(a)|(x)", r"(\d+):(\d+)", r"([a-z])([a-z])([a-z])", r"(é)", r"\b(\w{4,})\b", r"(a(b)?)+"]:
        want = [c.to_list() for c in ref.extract(p)]
        got = [lst(c) for c in dev.extract(p)]
        assert got == want, ("extract", p)
        wr = [r.to_list() for r in ref.extract_record(p)]
        gr = [lst(r) for r in dev.extract_record(p)]
        assert gr == wr, ("extract_record", p)
    assert dev.extract(r"\d+") == []  # no capture groups -> no columns (extract.cu:96-100)


def test_replace_with_backrefs(cols, oracle):
    strs, dev, ref = cols
    cases = [(r"([a-z])-([a-z])", r"X\1+\2Z"), (r"(\w)(\w)", r"\2\1"), (r"(\w+) (\w+)", r"[\2|\1]"), (r"([a-z])([0-9])", r"\0!"),
             (r"(\d+)", r"<\1>"), (r"(a)|(b)", r"<\1\2>"), (r"(\w+)@(\w+)", r"\2 at \1"), (r"(é)(.)", r"\2\1"), (r"(\w)(\w)?", r"\2-\3-\12"),
             (r"\b(\w)(\w*)\b", r"\2\1ay"), (r"(l+)", "none"), (r"(a(b)?)", r"[\2]")]
    for p, repl in cases:
        want = ref.replace_with_backrefs(p, repl).to_list()
        got = oracle.unpack(*dev.replace_with_backrefs(p, repl).to_arrays())
        assert got == want, (p, repl)


def test_rows_with_nul_bytes_fall_back_exactly(oracle):
    from custrings_b200 import nvstrings
    from custrings_b200._lib import lib
    strs = ["a\x00bcd efgh", "abcd\x00", "\x00abcd", "ab\x00", "\x00", "x\x00y\x00z abcd", "plain words here", None]
    dev, ref = nvstrings.to_device(strs), oracle.RefStrings.from_list(strs)
    for p in [r"c", r"\w+", r"\b\w{4,}\b", r"a\w+", r"[a-z]+", r"^a", r"d$", r"\W"]:
        assert [False if x is None else x for x in dev.contains(p)] == ref.contains_re(p)[0].tolist(), p
        assert [0 if x is None else x for x in dev.count(p)] == ref.count_re(p)[0].tolist(), p
        assert lib().custr_last_regex_tier() == b"pikevm"
        assert oracle.unpack(*dev.replace(p, "#").to_arrays()) == ref.replace_re(p, "#").to_list(), p
    clean = nvstrings.to_device(["abcd efgh", "x1 y22 z333"])
    assert clean.count(r"\d+") == [0, 3] and lib().custr_last_regex_tier() == b"bitcount"   # counted inside the chain kernel
    assert clean.count(r"x\d+") == [0, 1] and lib().custr_last_regex_tier() == b"bitspans"  # prefix outside the loop class: stream walk
    lib().custr_set_regex_tier(2)  # generic bitstream kernel forced: span streams unavailable -> scalar chain matcher
    try:
        assert clean.count(r"\d+") == [0, 3] and lib().custr_last_regex_tier() == b"chainspan"
    finally:
        lib().custr_set_regex_tier(0)




def test_count_replace_span_streams(oracle):
    """count_re / replace_re of last-loop chains through the chain kernel's bit streams (tier 'bitspans'), long and
    short rows, multi-byte characters, rows straddling windows and work items"""
    from custrings_b200 import nvstrings
    from custrings_b200._lib import lib
    rng = random.Random(31)
    strs = corpus.STRINGS + corpus.random_strings(rng, 3000)
    # long rows: many windows per row; plus a block of empty / null rows
    for k in range(40):
        strs.append(" ".join(rng.choice(["abcd", "x", "héllo", "12345", "a_b", "日本語", "wörld9", "_", "zz😀zz"]) for _ in range(rng.randrange(200, 900))))
    strs += ["", None, "", "a"] * 50
    rng.shuffle(strs)
    dev, ref = nvstrings.to_device(strs), oracle.RefStrings.from_list(strs)
    used = counted = 0
    for p in SPAN_PATTERNS:
        want = ref.count_re(p)[0].tolist()
        got = _none_to(0, dev.count(p))
        tier = lib().custr_last_regex_tier()
        assert got == want, (p, tier)
        used += tier in (b"bitspans", b"bitcount")
        counted += tier == b"bitcount"
        for repl, mx in (("<>", -1), ("", 2), ("é日", 1), ("#", 0), ("<" * 37 + ">", -1)):  # the long one overflows the splice tile
            want = ref.replace_re(p, repl, mx).to_list()
            assert oracle.unpack(*dev.replace(p, repl, mx).to_arrays()) == want, (p, repl, mx, lib().custr_last_regex_tier())
    assert used >= 15 and counted >= 8, (used, counted)


def test_large_programs_beyond_1024_instructions(oracle):
    """Programs with more instructions than the largest in-thread list tier run from the global scratch arena (the reference
    sizes global scratch per row for them, regexec.cpp:81-95): the 108 / 113-character patterns of the reference's
    python/tests/test_regex.py:256-273 and patterns of 1100-3000 instructions, all checked against the oracle."""
    from custrings_b200 import nvstrings
    base = "hello @abc @def world The quick brown @fox jumps over the lazy @dog hello http://www.world.com I'm here @home"
    rows = [base, "1234567890" * 11, "abcdefghijklmnopqrstuvwxyz" * 6, base + " zzzz", "", None, ("ab" * 700) + "c", "x" + "ab" * 1500 + "cd"]
    dev = nvstrings.to_device(rows)
    ref = oracle.RefStrings.from_list(rows)
    pats = [base, base + " zzzz", "ab" * 600, "(ab){550}c", "a?" * 400 + "b" * 400, "[a-c]{1200}", "ab" * 1400 + "c?d"]
    for p in pats:
        want, _ = ref.contains_re(p)
        assert dev.contains(p) == [None if r is None else bool(w) for r, w in zip(rows, want)], p[:40]
        wantm, _ = ref.match(p)
        assert dev.match(p) == [None if r is None else bool(w) for r, w in zip(rows, wantm)], p[:40]
    for p in ("ab" * 600, "(ab){550}"):
        wantc, _ = ref.count_re(p)
        assert dev.count(p) == [None if r is None else int(w) for r, w in zip(rows, wantc)], p[:40]
        assert oracle.unpack(*dev.replace(p, "#").to_arrays()) == oracle.unpack(*ref.replace_re(p, "#").to_arrays())


def test_runtime_compiled_plan_kernels(cols):
    """regex_jit.cu: the chain kernel compiled at run time (NVRTC) with every flag / class atom of the plan as a literal must give
    exactly what the ahead-of-time kernels and the oracle give — literals, classes with ranges and negation, assertions,
    optional steps / early exits, anchored search (match), multi-class chains."""
    from custrings_b200._lib import lib
    L = lib()
    pats = [r"\b\w{4,}\b", r"\d+", "Sun", r"[a-f]{3}\b", r"^\w{8}", r"z\w*$", r"colou?r", r"\d{1,3}", r"[^a-z ]+", r"q[aeiou]\w+", "é",
            r"\s[A-Z]\w", r"war(n|ning)?", r"\bthe\b", r"a.c", r"x?y?z"]
    before = L.custr_jit_launch_count()
    served = 0
    strs, dev, ref = cols
    try:
        for p in pats:
            L.custr_set_jit(0, 0)
            want_c, want_m = dev.contains(p), dev.match(p)
            L.custr_set_jit(2, 0)
            n0 = L.custr_jit_launch_count()
            got_c, got_m = dev.contains(p), dev.match(p)
            served += L.custr_jit_launch_count() > n0
            assert got_c == want_c and got_m == want_m, (p, L.custr_jit_note())
            rc, _ = ref.contains_re(p)
            assert got_c == [None if g is None else bool(w) for g, w in zip(got_c, rc)], p
    finally:
        L.custr_set_jit(1, 0)
    # (patterns that do not lower to a chain — nested alternations — never reach the chain kernels)
    assert L.custr_jit_launch_count() > before and served >= 10, (served, L.custr_jit_note().decode())

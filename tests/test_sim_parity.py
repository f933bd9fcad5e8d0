"""CPU-only parity of the engine LOGIC: the same __host__ __device__ per-row code the CUDA kernels run
(regex_vm.cuh) and the host executor of the bitstream plan are driven on the host by tests/sim and compared with the
oracle (the reference compiled for the CPU).  The GPU parity proper lives in the `-m gpu` tests."""
import random

import numpy as np
import pytest

from tests import corpus


@pytest.fixture(scope="module")
def data(oracle):
    from tests import simlib
    simlib.lib()
    rng = random.Random(99)
    strs = corpus.STRINGS + corpus.random_strings(rng, 250)
    chars, offsets, validity, nulls = oracle.pack(strs)
    return strs, chars, offsets, validity, oracle.RefStrings.from_list(strs)


SAFE_PATTERNS = [p for p in corpus.PATTERNS if p not in (r"(a|b)*c", r"((a|b)c)*d", "a+*")]


def test_vm_contains_match_count(data):
    from tests import simlib
    strs, chars, offsets, validity, ref = data
    for p in SAFE_PATTERNS + corpus.random_patterns(21, 80):
        for name, refv, simv in (("contains", ref.contains_re(p), simlib.bool_search(chars, offsets, validity, p, False)),
                                 ("match", ref.match(p), simlib.bool_search(chars, offsets, validity, p, True)),
                                 ("count", ref.count_re(p), simlib.count(chars, offsets, validity, p))):
            assert np.array_equal(refv[0], simv[0]) and refv[1] == simv[1], (name, p)


def test_vm_replace_re(data, oracle):
    from tests import simlib
    strs, chars, offsets, validity, ref = data
    for p in [r"\b\w{4,}\b", r"\d+", "a*", r"\s", "é", r"[^a-c]+", r"\bthe\b|\bfox\b", "x*?y", "^", "$", r"\b"] + corpus.random_patterns(22, 40):
        for repl, mx in (("<>", -1), ("", 2)):
            want = ref.replace_re(p, repl, mx).to_arrays()
            got = simlib.replace_re(chars, offsets, validity, p, repl, mx)
            assert np.array_equal(want[0], got[0]) and np.array_equal(want[1], got[1]), (p, repl, mx)


def test_vm_replace_re_multi(data, oracle):
    from tests import simlib
    strs, chars, offsets, validity, ref = data
    pats = [r"\d+", "[tT]he", r"\s+", "é"]
    for repls in (["#", "THE", "_", "e"], ["."]):
        rc, ro, rv, _ = oracle.pack(repls)
        want = ref.replace_re_multi(pats, oracle.RefStrings.from_list(repls)).to_arrays()
        got = simlib.replace_re_multi(chars, offsets, validity, pats, rc, ro, rv)
        assert np.array_equal(want[0], got[0]) and np.array_equal(want[1], got[1])


def test_bitstream_lowering_equals_reference(data):
    """every pattern the lowering accepts must give the reference's answer (ASCII rows via the plan, the rest via the VM)"""
    from tests import simlib
    strs, chars, offsets, validity, ref = data
    eligible = 0
    for p in SAFE_PATTERNS + corpus.random_patterns(23, 120):
        for anchored in (False, True):
            got, cnt = simlib.bits_bool(chars, offsets, validity, p, anchored)
            if got is None:
                continue
            eligible += 1
            want, wcnt = ref.match(p) if anchored else ref.contains_re(p)
            assert np.array_equal(want, got) and wcnt == cnt, (p, anchored)
    assert eligible > 100


def test_chain_model_equals_reference(data):
    """the chain model the 64-bit kernel evaluates (optional steps, early exits, up to 8 classes) gives the reference's
    answer for every pattern the lowering turns into a chain"""
    from tests import simlib
    strs, chars, offsets, validity, ref = data
    wide = [r"colou?r", r"https?://", r"warning", r"abcdefgh", r"qu?ick\s+brown", r"\d{1,3}", r"\d{2,4}:\d\d?", r"ab*c", r"a?b?c",
            r"x\d{2,4}\b", r"la+zy dog", r"[Tt]he quick", r"\w+\d?", r"\bthe\b", r"^\w+ ?", r"o?ver$", r"\Bx?y*z"]
    chains = optional = 0
    for p in SAFE_PATTERNS + wide + corpus.random_patterns(29, 200):
        for anchored in (False, True):
            got, cnt = simlib.chain_bool(chars, offsets, validity, p, anchored)
            if got is None:
                continue
            chains += 1
            optional += "?" in p or "*" in p or "," in p
            want, wcnt = ref.match(p) if anchored else ref.contains_re(p)
            assert np.array_equal(want, got) and wcnt == cnt, (p, anchored)
    assert chains > 80 and optional > 20, (chains, optional)


def test_headline_pattern_is_bitstream_eligible():
    from tests import simlib
    d = simlib.describe(r"\b\w{4,}\b")
    assert "bitstream: chain classes=1 steps=4" in d and "WORD" in d
    assert "not eligible" in simlib.describe(r"a*")       # nullable -> exact VM
    assert "not eligible" in simlib.describe(r"(ab)+c")   # loop over two instructions -> exact VM


def test_compiler_quirks():
    """documented lexer quirks of the reference"""
    from tests import simlib
    assert "CHAR 0x40" in simlib.describe(r"\x4a")        # hex digit 'a' is dropped (regcomp.cpp:362-366): 0x40 + 0
    assert "CHAR 0x4b" in simlib.describe(r"\x4b")
    assert simlib.describe(r"\101x").count("CHAR") == 1    # octal escape swallows the following char
    d = simlib.describe(r"(?:ab){2}")
    assert d.count("CHAR 0x61") == 2 and "LBRA" not in d
    assert simlib.describe(r"(ab){2}").count("LBRA") == 2  # capture groups are duplicated by {n}


def test_chain_span_fast_path_equals_reference(data, oracle):
    """count_re / replace_re through the scalar leftmost-longest chain matcher (chain_spans.cuh) for every eligible
    pattern (last-loop greedy chains)"""
    from tests import simlib
    strs, chars, offsets, validity, ref = data
    eligible = 0
    extra = [r"\b\w{4,}\b", r"\d+", r"[a-z]+\b", r"^\w+", r"\w+$", r"é+", r"\s+", r"a\w+", r"\Bb+", r"\d{2}:", r"[^a]+", r".+", r"x\b", r"_+\b"]
    for p in SAFE_PATTERNS + extra + corpus.random_patterns(24, 200):
        got, cnt = simlib.chain_count(chars, offsets, validity, p)
        if got is None:
            continue
        eligible += 1
        want, wcnt = ref.count_re(p)
        assert np.array_equal(want, got) and wcnt == cnt, p
        for repl, mx in (("<>", -1), ("", 2)):
            w = ref.replace_re(p, repl, mx).to_arrays()
            g = simlib.chain_replace(chars, offsets, validity, p, repl, mx)
            assert np.array_equal(w[0], g[0]) and np.array_equal(w[1], g[1]), (p, repl, mx)
    assert eligible > 40
    assert simlib.chain_count(chars, offsets, validity, r"[^\w]+?")[0] is None   # lazy loop: not eligible
    assert simlib.chain_count(chars, offsets, validity, r"\d+:\d+")[0] is None   # loop before the last step

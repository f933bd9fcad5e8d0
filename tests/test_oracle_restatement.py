"""Pins oracle/restate.py (plain-Python restatement of the reference algorithms) against the real reference
(oracle/_ref) on small seeded inputs, and checks this repo's regex compiler against the reference compiler's
program dump.  CPU only."""
import random

import pytest

from tests import corpus
from tests.adapters import OracleAPI
from oracle import restate


def _enc(strs):
    return [None if s is None else s.encode("utf-8") for s in strs]


def _dec(strs):
    return [None if s is None else s.decode("utf-8") for s in strs]


@pytest.fixture(scope="module")
def data(oracle):
    rng = random.Random(5)
    strs = _enc(list(corpus.STRINGS) + corpus.random_strings(rng, 50))
    api = OracleAPI(oracle)
    return api, api.column(strs), strs


def test_vm_restatement_matches_reference(data):
    api, col, strs = data
    pats = list(corpus.PATTERNS) + corpus.random_patterns(11, 60)
    null = [i for i, s in enumerate(strs) if s is None]
    for p in pats:
        want = [bool(x) for x in api.contains_re(col, p)]
        assert restate.contains_re(strs, p) == want, p
        want = [bool(x) for x in api.match(col, p)]
        assert restate.contains_re(strs, p, anchored=True) == want, p
        want = api.count_re(col, p)
        got = restate.count_re(strs, p)
        for i in null:  # reference leaves null rows at its own sentinel
            got[i] = want[i]
        assert got == want, p
        assert _dec(restate.replace_re(strs, p, b"<>")) == api.replace_re(col, p, "<>"), p
        assert _dec(restate.replace_re(strs, p, b"", 1)) == api.replace_re(col, p, "", 1), p


def test_literal_ops_restatement_matches_reference(data):
    api, col, strs = data
    for sub in ["a", "ab", "é", "日", " ", "xyz", "."]:
        b = sub.encode()
        for (s, e) in [(0, -1), (1, -1), (2, 5), (0, 3), (3, 2)]:
            assert restate.find(strs, b, s, e) == api.find(col, sub, s, e), (sub, s, e)
            assert restate.find(strs, b, s, e, reverse=True) == api.rfind(col, sub, s, e), (sub, s, e)
        for n in (-1, 1, 2):
            assert _dec(restate.replace(strs, b, b"#", n)) == api.replace(col, sub, "#", n), (sub, n)
    assert restate.hash_(strs) == api.hash(col)


def test_split_tokenize_category_restatement_matches_reference(data):
    api, col, strs = data
    for delim in [None, " ", ",", "ab", "é"]:
        for n in (-1, 1, 2):
            want = api.split(col, delim, n)
            got = [_dec(c) for c in restate.split(strs, None if delim is None else delim.encode(), n)]
            assert got == want, (delim, n)
    for delim in [None, " ", ",", "é"]:
        assert _dec(restate.tokenize(strs, None if delim is None else delim.encode())) == api.tokenize(col, delim), delim
    keys, vals = restate.category(strs)
    rk, rv = api.category(col)
    assert _dec(keys) == rk and vals == list(rv)


# ---- this repo's compiler vs the reference compiler's program ------------------------------------------------------
def _canon(insts, classes, start, starts):
    """Graph walk from the start instruction assigning ids in visit order, so two programs compare equal when they are
    isomorphic (instruction numbering differs once NOPs are stripped in a different order)."""
    ids, order, stack = {}, [], [start]
    while stack:
        i = stack.pop()
        if i in ids:
            continue
        ids[i] = len(ids)
        order.append(i)
        t, u1, u2 = insts[i]
        if t == restate.OR:
            stack.extend([u1, u2])   # visits left (u2) first
        elif t != restate.END:
            stack.append(u2)
    known = (restate.CHAR, restate.RBRA, restate.LBRA, restate.OR, restate.ANY, restate.ANYNL, restate.BOL, restate.EOL,
             restate.CCLASS, restate.NCCLASS, restate.BOW, restate.NBOW, restate.END)
    out = []
    for i in order:
        t, u1, u2 = insts[i]
        if t not in known:   # malformed construct: the reference leaves an untyped instruction, this repo OP_BAD
            out.append(("bad", ids[u2]))
        elif t == restate.OR:
            out.append((t, ids[u1], ids[u2]))
        elif t == restate.END:
            out.append((t,))
        elif t in (restate.CCLASS, restate.NCCLASS):
            out.append((t, classes[u1][0], tuple(classes[u1][1]), ids[u2]))
        else:
            out.append((t, u1 & 0xFFFFFFFF, ids[u2]))
    return out, [ids[s] for s in starts]


def test_compiler_topology_matches_reference(oracle):
    from tests import simlib
    if not simlib.available():
        pytest.skip("host simulation library not built")
    pats = list(corpus.PATTERNS) + corpus.random_patterns(23, 200) + [r"\b\w{4,}\b", r"a{2,3}b|c", r"(\d+)-(\d+)", r"[^a-c\d]x", r"\101x", r"\x41b"]
    for p in pats:
        ref = restate.reference_program(p)
        mine = simlib.program_dump(p)
        a = _canon(ref["insts"], ref["classes"], ref["start"], ref["starts"])
        b = _canon(mine["insts"], mine["classes"], mine["start"], mine["starts"])
        assert a == b, p
        assert ref["groups"] == mine["groups"], p

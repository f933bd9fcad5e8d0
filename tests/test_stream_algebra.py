"""The bit-stream formulations behind the splice kernels, executed position by position in plain Python and checked against
Python's own string functions / `re` — an executable statement of the algebra the CUDA code implements (split_bits.cuh,
replace_bits.cuh MODE 0 and MODE 1), independent of it.  The GPU tests check the kernels against the oracle; these check that
the formulation itself is the reference's per-row semantics (custring_view.inl:1004-1060,1169-1279; replace.cu:39-107).
No GPU, no library."""
import random
import re


def _alnum(c):
    return c.isalnum() and ord(c) < 128


def _word(c):
    return _alnum(c) or c == "_"


# alnum-based word boundary of the reference (regexec.inl BOW / NBOW), as a lookaround for Python's re
_B = r"(?:(?<![0-9A-Za-z])(?=[0-9A-Za-z])|(?<=[0-9A-Za-z])(?![0-9A-Za-z]))"


def _flatten(rows):
    """chars of the non-null rows back to back + ROWSTART positions (one per row, null rows have no bytes)"""
    chars, starts = "", []
    for r in rows:
        starts.append(len(chars))
        chars += r or ""
    return chars, starts


def splice_single_class(row, n, lead_b, trail_b, cls, repl):
    """replace_bits.cuh MODE 1 on one row: S = spread(M, K), FM = first M of a run, R = A spread backwards through K,
    occurrence starts FM & R, DROP = (S & R) | the n-1 characters in front of every occurrence start."""
    L = len(row)
    C = [cls(c) for c in row]

    def bnd(i):
        a = _alnum(row[i - 1]) if i > 0 else False
        b = _alnum(row[i]) if i < L else False
        return a != b

    K = [C[p] and p != 0 for p in range(L)]  # the match may continue INTO this byte (cleared at the row start)
    M = [False] * L
    for p in range(L):
        s = p - (n - 1)
        if s >= 0 and all(C[s:p + 1]) and (not lead_b or bnd(s)):
            M[p] = True
    A = [(not trail_b or bnd(p + 1)) for p in range(L)]
    S = [False] * L
    for p in range(L):
        S[p] = M[p] or (p > 0 and S[p - 1] and K[p])
    FM = [M[p] and not (p > 0 and S[p - 1] and K[p]) for p in range(L)]
    R = [False] * L
    for p in range(L - 1, -1, -1):
        R[p] = A[p] or (p + 1 < L and R[p + 1] and K[p + 1])
    FMv = [FM[p] and R[p] for p in range(L)]
    D = [S[p] and R[p] for p in range(L)]
    for p in range(L):
        if FMv[p]:
            for q in range(p - (n - 1), p):
                D[q] = True
    out = []
    for p in range(L):
        if FMv[p]:
            out.append(repl)
        if not D[p]:
            out.append(row[p])
    return "".join(out)


def test_replace_re_single_class_chain_algebra():
    rng = random.Random(3)
    for _ in range(6000):
        row = "".join(rng.choice("ab1_ ,") for _ in range(rng.randint(0, 14)))
        n = rng.randint(1, 4)
        lb, tb = rng.random() < .5, rng.random() < .5
        cls, cre = ((_word, r"[0-9A-Za-z_]"), (lambda c: c in "ab", "[ab]"))[rng.randrange(2)]
        pat = (_B if lb else "") + cre + "{%d,}" % n + (_B if tb else "")
        assert splice_single_class(row, n, lb, tb, cls, "#") == re.sub(pat, "#", row), (row, pat)


def splice_literal(rows, target, repl):
    """replace_bits.cuh MODE 0 on a column: M = E_0 & (E_1 >> 1) & .. with the shift-downs stopping at ROWSTART, DROP = M smeared
    up by m-1, output = kept bytes with repl spliced in at every M bit, new offsets = outputs before each row start."""
    chars, starts = _flatten(rows)
    L, m = len(chars), len(target)
    RS = [False] * (L + 1)
    for s in starts:
        RS[s] = True
    X = [chars[p] == target[m - 1] for p in range(L)]
    for k in range(m - 2, -1, -1):
        X = [(p + 1 < L and X[p + 1] and not RS[p + 1]) and chars[p] == target[k] for p in range(L)]
    D = [False] * L
    for p in range(L):
        if X[p]:
            for q in range(p, p + m):
                D[q] = True
    out, offs = [], []
    produced = 0
    nxt = 0
    for p in range(L + 1):
        while nxt < len(starts) and starts[nxt] == p:
            offs.append(produced)
            nxt += 1
        if p == L:
            break
        if X[p]:
            out.append(repl)
            produced += len(repl)
        if not D[p]:
            out.append(chars[p])
            produced += 1
    offs.append(produced)
    flat = "".join(out)
    return [None if r is None else flat[offs[i]:offs[i + 1]] for i, r in enumerate(rows)]


def test_literal_replace_algebra_border_free_targets():
    rng = random.Random(5)
    for _ in range(1500):
        rows = [None if rng.random() < .1 else "".join(rng.choice("abc ,") for _ in range(rng.randint(0, 12))) for _ in range(rng.randint(1, 8))]
        for target in ("ab", "abc", " ", "c,", "b"):
            repl = rng.choice(["", "X", "abab", target + target])
            want = [None if r is None else r.replace(target, repl) for r in rows]
            assert splice_literal(rows, target, repl) == want, (rows, target, repl)


def split_events(rows, delim):
    """split_bits.cuh on a column without empty valid rows: events = first byte of every valid row, then every delimiter byte;
    token offset = non-delimiter bytes before the event; row_offsets[r] = events before the row's first byte."""
    chars, starts = _flatten(rows)
    valid_start = {s for s, r in zip(starts, rows) if r}
    T_before, events, ev_before = 0, [], []
    for p, c in enumerate(chars):
        ev_before.append(len(events))
        if p in valid_start:
            events.append(T_before)
        if c == delim:
            events.append(T_before)
        else:
            T_before += 1
    ev_before.append(len(events))
    kept = "".join(c for c in chars if c != delim)
    events.append(len(kept))
    row_off = [ev_before[s] for s in starts] + [len(events) - 1]
    return [None if r is None else [kept[events[t]:events[t + 1]] for t in range(row_off[i], row_off[i + 1])] for i, r in enumerate(rows)]


def test_split_record_event_algebra():
    rng = random.Random(7)
    for _ in range(1500):
        rows = [None if rng.random() < .15 else "".join(rng.choice("ab, ") for _ in range(rng.randint(1, 10))) for _ in range(rng.randint(1, 8))]
        for delim in (",", " ", "a"):
            want = [None if r is None else r.split(delim) for r in rows]
            assert split_events(rows, delim) == want, (rows, delim)

"""CPU restatement of the reference's algorithms for the hot path, in plain Python / numpy.

TEST INFRASTRUCTURE ONLY (importers: tests/).  It is deliberately independent of everything under custrings_b200/:

* the regex VM below re-states `dreprog::regexec` (cpp/src/regex/regexec.inl:204-442) character by character and runs the
  program compiled by the REFERENCE's own host compiler (obtained through oracle/_ref via `ref_regex_dump`), so it checks
  the semantics of the Pike VM without depending on this repo's compiler or kernels;
* the row drivers re-state count.cu:43-55,174-195 and replace.cu:50-106;
* find / split / tokenize / category / hash re-state custring_view.inl:481-610,1169-1279, split.cu:768-806,892-941,
  text/tokens.cu:41-121, category/NVCategory.cu:250-298 + custring.inl:240-261, custring.inl:164-232.

The restatement itself is pinned against the real reference in tests/test_oracle_restatement.py (small cases: these are
pure-Python loops).  Strings are `bytes` (UTF-8) or None.
"""
import ctypes as C
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# reference instruction types (regcomp.h:25-40)
CHAR, RBRA, LBRA, OR, ANY, ANYNL, BOL, EOL, CCLASS, NCCLASS, BOW, NBOW, END = 0o177, 0o201, 0o202, 0o204, 0o300, 0o301, 0o303, 0o304, 0o305, 0o306, 0o307, 0o310, 0o377

_flags = None


def unicode_flags():
    """65536-entry class table (same generator the product build uses; provenance in tools/gen_unicode_flags.py)"""
    global _flags
    if _flags is None:
        sys.path.insert(0, os.path.join(_ROOT, "tools"))
        import gen_unicode_flags
        _flags = bytes(gen_unicode_flags.build_table())
    return _flags


def reference_program(pattern):
    """Program compiled by the reference's regcomp.cpp, as plain Python data."""
    from oracle import ref
    L = ref.lib()
    out = np.zeros(1 << 16, np.int32)
    pat = pattern.encode("utf-8") if isinstance(pattern, str) else pattern
    n = L.ref_regex_dump(pat, out.ctypes.data_as(C.c_void_p), len(out))
    w = out[:n].tolist()
    ninsts, start, ngroups, nstarts, nclasses = w[:5]
    p = 5
    insts = [tuple(w[p + 3 * i: p + 3 * i + 3]) for i in range(ninsts)]   # (type, u1, u2)
    p += 3 * ninsts
    starts = w[p:p + nstarts]
    p += nstarts
    classes = []
    for _ in range(nclasses):
        builtins, cnt = w[p], w[p + 1]
        classes.append((builtins, [x & 0xFFFFFFFF for x in w[p + 2:p + 2 + cnt]]))
        p += 2 + cnt
    return {"insts": insts, "start": start, "groups": ngroups, "starts": starts, "classes": classes}


# ---- characters: UTF-8 bytes packed big-endian in an int (custring_view.inl:1724-1744), code point (util.inl:51-75) -------
def to_chars(b):
    out, i = [], 0
    while i < len(b):
        c = b[i]
        w = 1 + (c & 0xF0 == 0xF0) + (c & 0xE0 == 0xE0) + (c & 0xC0 == 0xC0)
        v = 0
        for k in range(w):
            v = (v << 8) | (b[i + k] if i + k < len(b) else 0)
        out.append(v)
        i += w
    return out


def char_bytes(ch):
    n = 1 + (ch > 0xFF) + (ch > 0xFFFF) + (ch > 0xFFFFFF)
    return ch.to_bytes(n, "big")


def codepoint(c):
    if c < 0x80:
        return c
    if c < 0xE000:
        return ((c & 0x1F00) >> 2) | (c & 0x3F)
    if c < 0xF00000:
        return ((c & 0x0F0000) >> 4) | ((c & 0x3F00) >> 2) | (c & 0x3F)
    if c <= 0xF8000000:
        return ((c & 0x03000000) >> 6) | ((c & 0x3F0000) >> 4) | ((c & 0x3F00) >> 2) | (c & 0x3F)
    return 0


def _alnum(c):
    cp = codepoint(c)
    return cp < 0x10000 and (unicode_flags()[cp] & 15) != 0


def class_match(cls, ch):  # regexec.inl:127-155
    builtins, chrs = cls
    for i in range(0, len(chrs), 2):
        if chrs[i] <= ch <= chrs[i + 1]:
            return True
    if not builtins:
        return False
    cp = codepoint(ch)
    if cp > 0xFFFF:
        return False
    fl = unicode_flags()[cp]
    alnum, space, digit = (fl & 15) != 0, (fl & 16) != 0, (fl & 4) != 0
    return bool((builtins & 1 and (ch == 0x5F or alnum)) or (builtins & 2 and space) or (builtins & 4 and digit) or
                (builtins & 8 and ch != 10 and ch != 0x5F and not alnum) or (builtins & 16 and not space) or
                (builtins & 32 and ch != 10 and not digit))


def regexec(prog, chars, begin, end):
    """regexec.inl:204-442 with groupId == 0; positions are character indices. Returns (match, begin, end)."""
    insts, classes = prog["insts"], prog["classes"]
    txtlen = len(chars)
    stype = insts[prog["start"]][0]
    starttype = stype if stype in (CHAR, BOL) else 0
    startchar = insts[prog["start"]][1] & 0xFFFFFFFF
    match, pos, eos = 0, begin, end
    mb = me = -1
    list1 = []  # (inst id, range.x)
    checkstart = starttype
    while True:
        if checkstart:
            if starttype == CHAR:
                if startchar == 0 or txtlen == 0 or pos > txtlen:
                    return match, mb, me
                try:
                    pos = chars.index(startchar, pos)
                except ValueError:
                    return match, mb, me
            elif starttype == BOL and pos != 0:
                if startchar != ord("^"):
                    return match, mb, me
                try:
                    pos = chars.index(10, pos - 1) + 1
                except ValueError:
                    return match, mb, me
        if (eos < 0 or pos < eos) and match == 0:
            have = {i for i, _ in list1}
            for sid in prog["starts"]:
                if sid not in have:
                    list1.append((sid, pos))
                    have.add(sid)
        c = chars[pos] if pos < txtlen else 0
        expanded = True
        guard = 0
        while expanded and guard <= len(insts) + 1:
            guard += 1
            expanded = False
            list2, seen = [], set()

            def act(i, x):
                if i not in seen:
                    seen.add(i)
                    list2.append((i, x))
            for iid, x in list1:
                t, u1, u2 = insts[iid]
                if t in (CHAR, ANY, ANYNL, CCLASS, NCCLASS, END):
                    act(iid, x)
                elif t in (LBRA, RBRA):
                    act(u2, x)
                    expanded = True
                elif t == BOL:
                    if pos == 0 or (u1 == ord("^") and chars[pos - 1] == 10):
                        act(u2, x)
                        expanded = True
                elif t == EOL:
                    if c == 0 or (u1 == ord("$") and c == 10):
                        act(u2, x)
                        expanded = True
                elif t in (BOW, NBOW):
                    cur = _alnum(c)
                    prev = _alnum(chars[pos - 1]) if pos else False
                    if (cur != prev) == (t == BOW):
                        act(u2, x)
                        expanded = True
                elif t == OR:
                    act(u1, x)
                    act(u2, x)
                    expanded = True
            list1 = list2
        list2, seen = [], set()
        for iid, x in list1:
            t, u1, u2 = insts[iid]
            go = False
            if t == CHAR:
                go = (u1 & 0xFFFFFFFF) == c
            elif t == ANY:
                go = c != 10
            elif t == ANYNL:
                go = True
            elif t == CCLASS:
                go = class_match(classes[u1], c)
            elif t == NCCLASS:
                go = not class_match(classes[u1], c)
            elif t == END:
                match, mb, me = 1, x, pos
                break
            if go and u2 not in seen:
                seen.add(u2)
                list2.append((u2, x))
        pos += 1
        list1 = list2
        checkstart = 0 if list1 else 1
        if not (c and (list1 or match == 0)):
            break
    return match, mb, me


def contains_re(strs, pattern, anchored=False):  # count.cu:43-55
    prog = reference_program(pattern)
    out = []
    for s in strs:
        if s is None:
            out.append(False)
            continue
        ch = to_chars(s)
        out.append(bool(regexec(prog, ch, 0, 1 if anchored else len(ch))[0]))
    return out


def count_re(strs, pattern):  # count.cu:174-195
    prog = reference_program(pattern)
    out = []
    for s in strs:
        n = 0
        if s is not None:
            ch = to_chars(s)
            begin = 0
            while begin <= len(ch):
                m, b, e = regexec(prog, ch, begin, len(ch))
                if not m:
                    break
                n += 1
                begin = e if e > b else b + 1
        out.append(n)
    return out


def replace_re(strs, pattern, repl, maxrepl=-1):  # replace.cu:50-106
    prog = reference_program(pattern)
    out = []
    for s in strs:
        if s is None:
            out.append(None)
            continue
        ch = to_chars(s)
        mxn = len(ch) if maxrepl < 0 else maxrepl
        pieces, lpos, begin = [], 0, 0
        while mxn > 0:
            m, b, e = regexec(prog, ch, begin, len(ch))
            if not m:
                break
            pieces.append(b"".join(char_bytes(x) for x in ch[lpos:b]) + repl)
            lpos, begin = e, e
            mxn -= 1
        out.append(b"".join(pieces) + b"".join(char_bytes(x) for x in ch[lpos:]))
    return out


# ---- literal ops ---------------------------------------------------------------------------------------------------
def _char_offset(s, bytepos):
    return sum(1 for b in s[:bytepos] if b & 0xC0 != 0x80)


def _byte_offset(s, chpos):
    off = 0
    for _ in range(chpos):
        if off >= len(s):
            break
        off += 1
        while off < len(s) and s[off] & 0xC0 == 0x80:
            off += 1
    return off


def find(strs, sub, start=0, end=-1, reverse=False):  # find.cu:75-120,163-199 / custring_view.inl:481-514,550-582
    out = []
    start = max(start, 0)
    for s in strs:
        if s is None:
            out.append(-2)
            continue
        if not sub:
            out.append(-1)
            continue
        nchars = _char_offset(s, len(s))
        count = end - start
        if count < 0 and not reverse:
            count = nchars
        e = start + count
        if e < 0 or e > nchars:
            e = nchars
        spos, epos = _byte_offset(s, start), _byte_offset(s, e)
        hit = s.rfind(sub, spos, epos) if reverse else s.find(sub, spos, epos)
        out.append(-1 if hit < 0 else _char_offset(s, hit))
    return out


def replace(strs, target, repl, maxrepl=-1):  # modify.cu:125-187
    out = []
    for s in strs:
        if s is None:
            out.append(None)
            continue
        mxn = _char_offset(s, len(s)) if maxrepl < 0 else maxrepl
        out.append(s.replace(target, repl, mxn))
    return out


def _split_row(s, d, limit):
    """custring_view.inl:1223-1248 (token count: the scan resumes at CHAR position pos + delimiter BYTES, so adjacent
    multi-byte delimiters are under-counted) + split.cu:768-806 (column walk; once a search fails every later column
    repeats the remainder)"""
    if len(s) == 0 or not d:
        return [s]
    nchars = _char_offset(s, len(s))
    hits, p = 0, s.find(d)
    while p >= 0:
        hits += 1
        cp = _char_offset(s, p) + len(d)
        p = s.find(d, _byte_offset(s, cp)) if cp <= nchars else -1
    dcount = hits + 1
    if limit > 0 and dcount > limit:
        dcount = limit
    toks, spos, failed = [], 0, False
    for c in range(dcount):
        b, e = spos, len(s)
        if not failed and c < dcount - 1:
            hit = s.find(d, spos)
            if hit < 0:
                failed = True
            else:
                e, spos = hit, hit + len(d)
        toks.append(s[b:e] if b < e else b"")
    return toks


def split(strs, delim, maxsplit=-1):  # split.cu:734-822 (byte delimiter) / :863-956 (whitespace when delim is None)
    limit = maxsplit + 1 if maxsplit > 0 else 0
    rows = []
    for s in strs:
        if s is None:
            rows.append(None)
        elif delim is None:
            toks, i = [], 0
            while True:
                while i < len(s) and s[i] <= 32:
                    i += 1
                if i >= len(s):
                    break
                if limit and len(toks) + 1 == limit:
                    toks.append(s[i:])
                    break
                j = i
                while j < len(s) and s[j] > 32:
                    j += 1
                toks.append(s[i:j])
                i = j
            rows.append(toks if toks else [None])
        else:
            rows.append(_split_row(s, delim, limit))
    ncols = max([len(r) for r in rows if r is not None] + [0])
    if ncols == 0:
        return [[None] * len(strs)]
    return [[(r[c] if r is not None and c < len(r) else None) for r in rows] for c in range(ncols)]


def tokenize(strs, delims=None):  # tokens.cu:41-121
    out = []
    dset = None if delims is None else set(to_chars(delims))
    for s in strs:
        if s is None:
            continue
        tok = []
        for ch in to_chars(s):
            is_d = ch <= 32 if dset is None else ch in dset
            if is_d:
                if tok:
                    out.append(b"".join(char_bytes(x) for x in tok))
                    tok = []
            else:
                tok.append(ch)
        if tok:
            out.append(b"".join(char_bytes(x) for x in tok))
    return out


def category(strs):  # NVCategory.cu:250-298: keys sorted by unsigned bytes (null first), values = key index
    keys = sorted({s for s in strs if s is not None})
    if any(s is None for s in strs):
        keys = [None] + keys
    index = {k: i for i, k in enumerate(keys)}
    return keys, [index[s] for s in strs]


def murmur3_32(b, seed=31):  # custring.inl:164-232
    def rotl(x, r):
        return ((x << r) | (x >> (32 - r))) & 0xFFFFFFFF
    h = seed
    n = len(b) // 4
    for i in range(n):
        k = int.from_bytes(b[4 * i:4 * i + 4], "little")
        k = (k * 0xcc9e2d51) & 0xFFFFFFFF
        k = rotl(k, 15)
        k = (k * 0x1b873593) & 0xFFFFFFFF
        h ^= k
        h = rotl(h, 13)
        h = (h * 5 + 0xe6546b64) & 0xFFFFFFFF
    tail = b[4 * n:]
    k = 0
    if len(tail) == 3:
        k ^= tail[2] << 16
    if len(tail) >= 2:
        k ^= tail[1] << 8
    if len(tail) >= 1:
        k ^= tail[0]
        k = (k * 0xcc9e2d51) & 0xFFFFFFFF
        k = rotl(k, 15)
        k = (k * 0x1b873593) & 0xFFFFFFFF
        h ^= k
    h ^= len(b)
    h ^= h >> 16
    h = (h * 0x85ebca6b) & 0xFFFFFFFF
    h ^= h >> 13
    h = (h * 0xc2b2ae35) & 0xFFFFFFFF
    h ^= h >> 16
    return h


def hash_(strs):
    return [0 if s is None else murmur3_32(s) for s in strs]

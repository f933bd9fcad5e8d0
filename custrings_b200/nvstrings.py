"""nvstrings — host-side mirror of the reference Python shim (python/nvstrings.py) for the hot path.

Same function / method names, argument meaning, defaults (`regex=True` for replace / contains,
nvstrings.py:1460,1951) and error behaviour (ValueError from bad arguments, None entries for null rows in host
results, pystrings.cpp:2644-2664) as the reference, but it drives libcustr.so (hand-written sm_100a CUDA over an
Arrow-layout column) through the C-ABI in include/custr.h instead of pyniNVStrings.

Methods of the reference that are outside the hot path raise NotImplementedError.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import as_ptr, check_handle, check_rc, lib


def _enc(s):
    if s is None:
        return None
    return s.encode("utf-8") if isinstance(s, str) else bytes(s)


def _pack(strs):
    n = len(strs)
    enc = [None if s is None else (s.encode("utf-8") if isinstance(s, str) else bytes(s)) for s in strs]
    lens = np.fromiter((0 if e is None else len(e) for e in enc), dtype=np.int64, count=n)
    if lens.sum() > 0x7FFFFFFF:
        raise ValueError("nvstrings: more than 2 GiB of characters in one instance (int32 offsets)")
    offsets = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(lens, out=offsets[1:])
    chars = np.frombuffer(b"".join(e for e in enc if e is not None), dtype=np.uint8)
    valid = np.fromiter((e is not None for e in enc), dtype=bool, count=n)
    validity = np.packbits(valid, bitorder="little") if n else np.zeros(0, np.uint8)
    return chars, offsets, validity, int(n - valid.sum())


def to_device(strs):
    """Create nvstrings instance from list of Python strings (None = null).  reference nvstrings.py:4"""
    if strs is None:
        raise ValueError("to_device: strs is None")
    if isinstance(strs, str):
        strs = [strs]
    chars, offsets, validity, nulls = _pack(list(strs))
    if chars.size == 0:
        chars = np.zeros(1, np.uint8)
    h = lib().custr_create_from_offsets(as_ptr(chars), len(offsets) - 1, as_ptr(offsets), as_ptr(validity) if nulls else None, nulls, 0)
    return nvstrings(check_handle(h, "to_device"))


def from_offsets(sbuf, obuf, scount, nbuf=None, ncount=0, bdevmem=False):
    """Create nvstrings from chars + int32 offsets[scount+1] (+ Arrow validity bits).  reference nvstrings.py:103"""
    if sbuf is None or obuf is None:
        raise ValueError("from_offsets: sbuf and obuf are required")
    h = lib().custr_create_from_offsets(as_ptr(sbuf), int(scount), as_ptr(obuf), as_ptr(nbuf), int(ncount), 1 if bdevmem else 0)
    return nvstrings(check_handle(h, "from_offsets"))


def from_device_view(chars, offsets, scount, validity=None, ncount=0, keepalive=None):
    """Zero-copy adoption of device buffers (e.g. torch tensors); `keepalive` objects are referenced by the result."""
    h = lib().custr_adopt_device(as_ptr(chars), int(scount), as_ptr(offsets), as_ptr(validity), int(ncount))
    s = nvstrings(check_handle(h, "from_device_view"))
    s._keepalive = (chars, offsets, validity, keepalive)
    return s


def from_index(pairs, count, bdevmem=True, stype=0):
    """NVStrings::create_from_index (NVStrings.h:98): `pairs` = address of `count` (device pointer, byte length) entries
    laid out like std::pair<const char*, size_t> (int | numpy uint64[count, 2] | torch tensor); bdevmem: the pair ARRAY is
    in device memory.  stype: 0 none, 1 length, 2 name, 3 both.  This is how cuDF / nvtext hand string views in."""
    h = lib().custr_create_from_index(as_ptr(pairs), count, 1 if bdevmem else 0, stype)
    return nvstrings(check_handle(h, "from_index"))


def create_from_ipc(ipc_data):
    """Column exported by another process on the same GPU (nvstrings.get_ipc_data()); zero copy.  reference nvstrings.py
    create_from_ipc -> NVStrings::create_from_ipc (ipc_transfer.h)"""
    buf = C.create_string_buffer(bytes(ipc_data), len(ipc_data))
    return nvstrings(check_handle(lib().custr_ipc_import(buf), "create_from_ipc"))


def from_strings(*args):
    """Concatenate nvstrings instances into one.  reference nvstrings.py:27"""
    cols = []
    for a in args:
        cols.extend(a if isinstance(a, (list, tuple)) else [a])
    parts = [c.to_arrays() for c in cols]
    chars = np.concatenate([p[0] for p in parts]) if parts else np.zeros(0, np.uint8)
    offs = [np.zeros(1, np.int64)]
    base = 0
    valids = []
    for (ch, off, val), c in zip(parts, cols):
        offs.append(off[1:].astype(np.int64) + base)
        base += int(off[-1])
        valids.append(np.unpackbits(val, bitorder="little")[: c.size()] if c.size() else np.zeros(0, np.uint8))
    offsets = np.concatenate(offs).astype(np.int32)
    valid = np.concatenate(valids) if valids else np.zeros(0, np.uint8)
    n = len(offsets) - 1
    nulls = int(n - valid.sum())
    return from_offsets(chars if chars.size else np.zeros(1, np.uint8), offsets, n, np.packbits(valid, bitorder="little"), nulls)


def free(dstrs):
    """Force free of the device memory held by an instance.  reference nvstrings.py:363"""
    dstrs._release()


def bind_cpointer(cptr, own=True):
    """Wrap an existing custr_column* handle.  reference nvstrings.py:370"""
    if cptr:
        return nvstrings(cptr, own)
    return None


class nvstrings:
    """Immutable column of UTF-8 strings resident in GPU memory (handle on a custr_column)."""

    def __init__(self, cptr, own=True):
        self.m_cptr = cptr
        self._own = own
        self._keepalive = None

    def _release(self):
        if getattr(self, "m_cptr", None) and self._own:
            try:
                lib().custr_column_free(self.m_cptr)
            except Exception:
                pass
        self.m_cptr = 0

    def __del__(self):
        self._release()

    def __str__(self):
        return str(self.to_host())

    def __repr__(self):
        return "<nvstrings count={}>".format(self.size())

    def __len__(self):
        return self.size()

    def __iter__(self):
        raise TypeError("iterable not supported by nvstrings")

    def __getitem__(self, key):
        if isinstance(key, int):
            n = self.size()
            if key < 0:
                key += n
            return self.gather([key])
        if isinstance(key, slice):
            start, stop, step = key.indices(self.size())
            if step == 1:
                h = lib().custr_slice_rows(self.m_cptr, start, max(stop, start))
                return nvstrings(check_handle(h, "sublist"))
            return self.gather(list(range(start, stop, step)))
        if isinstance(key, (list, np.ndarray)):
            return self.gather(key)
        raise KeyError("key must be int, slice or list of ints")

    def get_cpointer(self):
        return self.m_cptr

    # ------------------------------------------------------------------ export / attributes
    def size(self):
        """Number of strings.  reference nvstrings.py:519"""
        return int(lib().custr_size(self.m_cptr))

    def to_arrays(self):
        """(chars uint8[], offsets int32[n+1], validity uint8[(n+7)//8]) numpy copies on the host."""
        n = self.size()
        total = int(lib().custr_chars_bytes(self.m_cptr))
        chars = np.zeros(max(total, 1), np.uint8)
        offsets = np.zeros(n + 1, np.int32)
        validity = np.zeros((n + 7) // 8, np.uint8)
        check_rc(lib().custr_create_offsets(self.m_cptr, as_ptr(chars), as_ptr(offsets), as_ptr(validity) if n else None, 0), "to_offsets")
        return chars[:total], offsets, validity

    def to_host(self):
        """List of Python strings, None for nulls.  reference nvstrings.py:464"""
        chars, offsets, validity = self.to_arrays()
        n = len(offsets) - 1
        valid = np.unpackbits(validity, bitorder="little")[:n].astype(bool) if n else np.zeros(0, bool)
        raw = chars.tobytes()
        return [raw[offsets[i]:offsets[i + 1]].decode("utf-8", "replace") if valid[i] else None for i in range(n)]

    def to_offsets(self, sbuf, obuf, nbuf=0, bdevmem=False):
        """Store chars / offsets / optional null bitmask into caller memory.  reference nvstrings.py:484"""
        return check_rc(lib().custr_create_offsets(self.m_cptr, as_ptr(sbuf), as_ptr(obuf), as_ptr(nbuf), 1 if bdevmem else 0), "to_offsets")

    def _valid_mask(self):
        n = self.size()
        bits = np.zeros((n + 7) // 8, np.uint8)
        if n:
            lib().custr_set_null_bitarray(self.m_cptr, as_ptr(bits), 0, 0)
        return np.unpackbits(bits, bitorder="little")[:n].astype(bool)

    def _host_result(self, fn, dtype, devptr, what, null_below=None, as_bool=False, *args):
        """Run fn(handle, *args, results, devmem) the way the n_* bindings do: devptr => in place on the device,
        else a Python list with None for null rows."""
        n = self.size()
        if devptr:
            check_rc(fn(self.m_cptr, *args, as_ptr(devptr), 1), what)
            return devptr
        out = np.zeros(max(n, 1), dtype)
        rc = check_rc(fn(self.m_cptr, *args, as_ptr(out), 0), what)
        if rc is not None and rc == -1 and n:
            return None
        out = out[:n]
        if null_below is not None:
            return [None if v < null_below else int(v) for v in out]
        valid = self._valid_mask()
        if as_bool:
            return [bool(v) if ok else None for v, ok in zip(out, valid)]
        return [int(v) if ok else None for v, ok in zip(out, valid)]

    def len(self, devptr=0):
        """Characters per string (None / -1 for null).  reference nvstrings.py:538"""
        return self._host_result(lib().custr_len, np.int32, devptr, "len", null_below=0)

    def byte_count(self, vals=0, bdevmem=False):
        """Bytes per string; returns the total.  reference nvstrings.py:567"""
        return int(lib().custr_byte_count(self.m_cptr, as_ptr(vals), 1 if bdevmem else 0))

    def set_null_bitmask(self, nbuf, bdevmem=False):
        """Arrow validity bits into nbuf; returns the null count.  reference nvstrings.py:598"""
        return check_rc(lib().custr_set_null_bitarray(self.m_cptr, as_ptr(nbuf), 0, 1 if bdevmem else 0), "set_null_bitmask")

    def null_count(self, emptyisnull=False):
        """reference nvstrings.py:622"""
        n = self.size()
        if not emptyisnull:
            return int(lib().custr_null_count(self.m_cptr))
        bits = np.zeros((n + 7) // 8 + 1, np.uint8)
        return check_rc(lib().custr_set_null_bitarray(self.m_cptr, as_ptr(bits), 1, 0), "null_count") if n else 0

    def hash(self, devptr=0):
        """MurmurHash3_32 (seed 31) of each string.  reference nvstrings.py:675"""
        return self._host_result(lib().custr_hash, np.uint32, devptr, "hash")

    def gather(self, indexes, count=0):
        """reference nvstrings.py:2394"""
        if isinstance(indexes, (list, tuple, np.ndarray)):
            idx = np.ascontiguousarray(indexes, np.int32)
            h = lib().custr_gather(self.m_cptr, as_ptr(idx), len(idx), 0)
        else:
            h = lib().custr_gather(self.m_cptr, as_ptr(indexes), int(count), 1)
        return nvstrings(check_handle(h, "gather"))

    sublist = gather

    def get_ipc_data(self):
        """bytes to hand to nvstrings.create_from_ipc() in another process (this column must stay alive meanwhile).
        reference nvstrings.py get_ipc_data -> NVStrings::create_ipc_transfer"""
        buf = C.create_string_buffer(96)
        check_rc(lib().custr_ipc_export(self.m_cptr, buf), "get_ipc_data")
        return buf.raw

    def copy(self):
        return self[0:self.size()]

    # ------------------------------------------------------------------ regex
    def contains(self, pat, regex=True, devptr=0):
        """True where `pat` (regex by default) is found.  reference nvstrings.py:1951 -> count.cu:59 / find.cu:237"""
        if pat is None:
            raise ValueError("contains: pat is None")
        fn = lib().custr_contains_re if regex else lib().custr_contains
        return self._host_result(fn, np.uint8, devptr, "contains", None, True, _enc(pat))

    def match(self, pat, devptr=0):
        """True where the regex matches at the start of the string.  reference nvstrings.py:1980 -> count.cu:113"""
        if pat is None:
            raise ValueError("match: pat is None")
        return self._host_result(lib().custr_match, np.uint8, devptr, "match", None, True, _enc(pat))

    def count(self, pat, devptr=0):
        """Number of regex matches per string.  reference nvstrings.py:2033 -> count.cu:199"""
        if pat is None:
            raise ValueError("count: pat is None")
        return self._host_result(lib().custr_count_re, np.int32, devptr, "count", None, False, _enc(pat))

    def replace(self, pat, repl, n=-1, regex=True):
        """Replace `pat` (regex by default) with `repl`, at most n times.  reference nvstrings.py:1460"""
        fn = lib().custr_replace_re if regex else lib().custr_replace
        h = fn(self.m_cptr, _enc(pat), _enc(repl), int(n))
        return nvstrings(check_handle(h, "replace"))

    def replace_with_backrefs(self, pat, repl):
        """Replace every match of `pat` by `repl` with its \\1..\\N back-references filled in from the capture groups.
        reference nvstrings.py:1540 -> replace_backref.cu:122"""
        if pat is None or pat == "":
            raise ValueError("replace_with_backrefs: pattern cannot be null or empty")
        h = lib().custr_replace_with_backrefs(self.m_cptr, _enc(pat), _enc(repl))
        return nvstrings(check_handle(h, "replace_with_backrefs"))

    def replace_multi(self, pats, repls, regex=True):
        """reference nvstrings.py:1487 -> replace_multi.cu:110 (regex) / modify.cu:263 (literal)"""
        if isinstance(repls, str):
            repls = to_device([repls])
        elif isinstance(repls, (list, tuple)):
            repls = to_device(list(repls))
        if regex:
            if isinstance(pats, nvstrings):
                pats = pats.to_host()
            arr = (C.c_char_p * len(pats))(*[_enc(p) for p in pats])
            h = lib().custr_replace_re_multi(self.m_cptr, arr, len(pats), repls.m_cptr)
        else:
            if isinstance(pats, (list, tuple)):
                pats = to_device(list(pats))
            h = lib().custr_replace_multi(self.m_cptr, pats.m_cptr, repls.m_cptr)
        return nvstrings(check_handle(h, "replace_multi"))

    def findall(self, pat):
        """All non-overlapping matches, column-major: result[c] holds the c-th match of every row (None where a row has
        fewer).  reference nvstrings.py:1921 -> findall.cu:99"""
        return self._columns(lib().custr_findall, pat, "findall")

    def findall_record(self, pat):
        """One nvstrings of matches per row (None for null rows).  reference nvstrings.py:1891 -> findall_record.cu:97"""
        rows = self.size()
        row_off = np.zeros(rows + 1, np.int32)
        tok = C.c_void_p()
        check_rc(lib().custr_findall_record(self.m_cptr, _enc(pat), C.byref(tok), as_ptr(row_off), 0), "findall_record")
        tokens = nvstrings(check_handle(tok.value, "findall_record"))
        # the reference returns an EMPTY instance (not None) for rows without matches, null rows included
        return [nvstrings(check_handle(lib().custr_slice_rows(tokens.m_cptr, int(row_off[i]), int(row_off[i + 1])), "findall_record"))
                for i in range(rows)]

    def extract(self, pat):
        """One nvstrings per capture group of the first match.  reference nvstrings.py:2127 -> extract.cu:69"""
        return self._columns(lib().custr_extract, pat, "extract")

    def extract_record(self, pat):
        """Per row: an nvstrings with one entry per capture group (None for null rows).  reference nvstrings.py:2097"""
        cols = [c.to_host() for c in self.extract(pat)]
        hit = self.contains(pat)
        # reference extract_record.cu: a matching row gets "" for a group without a span, a non-matching row all-null
        return [to_device([(c[i] if c[i] is not None else ("" if hit[i] else None)) for c in cols]) for i in range(self.size())]

    def _columns(self, fn, pat, what):
        if pat is None:
            raise ValueError(what + ": pat is None")
        cap = 64
        while True:
            out = (C.c_void_p * cap)()
            cols = check_rc(fn(self.m_cptr, _enc(pat), out, cap), what)
            got = [nvstrings(out[i]) for i in range(min(max(cols, 0), cap))]
            if cols <= cap:
                return got
            del got
            cap = cols

    # ------------------------------------------------------------------ literal find
    def find(self, sub, start=0, end=None, devptr=0):
        """Character position of the first `sub` in [start,end), -1 if absent.  reference nvstrings.py:1796"""
        return self._host_result(lib().custr_find, np.int32, devptr, "find", -1, False, _enc(sub), int(start), -1 if end is None else int(end))

    def rfind(self, sub, start=0, end=None, devptr=0):
        """reference nvstrings.py:1861"""
        return self._host_result(lib().custr_rfind, np.int32, devptr, "rfind", -1, False, _enc(sub), int(start), -1 if end is None else int(end))

    def startswith(self, pat, devptr=0):
        """reference nvstrings.py:2049"""
        return self._host_result(lib().custr_startswith, np.uint8, devptr, "startswith", None, True, _enc(pat))

    def _is(self, kind, devptr, what):
        return self._host_result(lib().custr_is_class, np.uint8, devptr, what, None, True, kind)

    def isalnum(self, devptr=0):
        """every character alphanumeric (at least one).  reference nvstrings.py isalnum -> attrs.cu:115"""
        return self._is(0, devptr, "isalnum")

    def isalpha(self, devptr=0):
        return self._is(1, devptr, "isalpha")

    def isdigit(self, devptr=0):
        return self._is(2, devptr, "isdigit")

    def isspace(self, devptr=0):
        return self._is(3, devptr, "isspace")

    def isdecimal(self, devptr=0):
        return self._is(4, devptr, "isdecimal")

    def isnumeric(self, devptr=0):
        return self._is(5, devptr, "isnumeric")

    def islower(self, devptr=0):
        return self._is(6, devptr, "islower")

    def isupper(self, devptr=0):
        return self._is(7, devptr, "isupper")

    def is_empty(self, devptr=0):
        """null rows are empty too.  attrs.cu:412"""
        if devptr:
            check_rc(lib().custr_is_class(self.m_cptr, 8, as_ptr(devptr), 1), "is_empty")
            return devptr
        out = np.zeros(max(self.size(), 1), np.uint8)
        check_rc(lib().custr_is_class(self.m_cptr, 8, as_ptr(out), 0), "is_empty")
        return [bool(v) for v in out[: self.size()]]

    def lower(self):
        """reference nvstrings.py lower -> case.cu:30"""
        return nvstrings(check_handle(lib().custr_case(self.m_cptr, 0), "lower"))

    def upper(self):
        return nvstrings(check_handle(lib().custr_case(self.m_cptr, 1), "upper"))

    def strip(self, to_strip=None):
        """reference nvstrings.py strip -> strip.cu:87; None strips space, newline and tab"""
        return nvstrings(check_handle(lib().custr_strip(self.m_cptr, _enc(to_strip) if to_strip is not None else None, 0), "strip"))

    def lstrip(self, to_strip=None):
        return nvstrings(check_handle(lib().custr_strip(self.m_cptr, _enc(to_strip) if to_strip is not None else None, 1), "lstrip"))

    def rstrip(self, to_strip=None):
        return nvstrings(check_handle(lib().custr_strip(self.m_cptr, _enc(to_strip) if to_strip is not None else None, 2), "rstrip"))

    def slice(self, start, stop=None, step=None):
        """characters [start, stop) of every string, every step-th one.  reference nvstrings.py slice -> substr.cu:39"""
        return nvstrings(check_handle(lib().custr_slice(self.m_cptr, int(start), -1 if stop is None else int(stop), 1 if step is None else int(step)), "slice"))

    def get(self, i):
        """the character at position i.  substr.cu:32"""
        return self.slice(i, i + 1, 1)

    def endswith(self, pat, devptr=0):
        """reference nvstrings.py:2073"""
        return self._host_result(lib().custr_endswith, np.uint8, devptr, "endswith", None, True, _enc(pat))

    def find_from(self, sub, starts=None, ends=None, devptr=0):
        """find() with per-row start / end character positions (host int32 sequences here; device pointers together with
        devptr).  reference nvstrings.py:1829 -> find.cu:123-160"""
        n = self.size()
        if devptr:
            check_rc(lib().custr_find_from(self.m_cptr, _enc(sub), as_ptr(starts) if starts else None, as_ptr(ends) if ends else None,
                                           as_ptr(devptr), 1), "find_from")
            return devptr
        hs = None if starts is None else np.ascontiguousarray(starts, np.int32)
        he = None if ends is None else np.ascontiguousarray(ends, np.int32)
        out = np.zeros(max(n, 1), np.int32)
        check_rc(lib().custr_find_from(self.m_cptr, _enc(sub), None if hs is None else as_ptr(hs), None if he is None else as_ptr(he),
                                       as_ptr(out), 0), "find_from")
        return [None if v < -1 else int(v) for v in out[:n]]

    def match_strings(self, strs, devptr=0):
        """Row-wise equality with another nvstrings of the same size.  reference nvstrings.py:2018 -> find.cu:276-313"""
        n = self.size()
        if devptr:
            check_rc(lib().custr_match_strings(self.m_cptr, strs.m_cptr, as_ptr(devptr), 1), "match_strings")
            return devptr
        out = np.zeros(max(n, 1), np.uint8)
        check_rc(lib().custr_match_strings(self.m_cptr, strs.m_cptr, as_ptr(out), 0), "match_strings")
        return [bool(v) for v in out[:n]]

    def find_multiple(self, strs, devptr=0):
        """reference nvstrings.py:2550: one row of positions per string"""
        n, m = self.size(), strs.size()
        if devptr:
            check_rc(lib().custr_find_multiple(self.m_cptr, strs.m_cptr, as_ptr(devptr), 1), "find_multiple")
            return devptr
        out = np.zeros(max(n * m, 1), np.int32)
        check_rc(lib().custr_find_multiple(self.m_cptr, strs.m_cptr, as_ptr(out), 0), "find_multiple")
        return out[: n * m].reshape(n, m).tolist()

    # ------------------------------------------------------------------ split
    def _split(self, fn, delimiter, n, what):
        cap = 64
        while True:
            out = (C.c_void_p * cap)()
            cols = check_rc(fn(self.m_cptr, _enc(delimiter), int(n), out, cap), what)
            if cols is not None and cols == -1:
                raise ValueError(what + ": " + _lib.last_error())
            got = [nvstrings(out[i]) for i in range(min(cols, cap))]
            if cols <= cap:
                return got
            del got
            cap = cols

    def split(self, delimiter=None, n=-1):
        """Column-major split: list of nvstrings, one per output column.  reference nvstrings.py:1069 -> split.cu:734,863"""
        return self._split(lib().custr_split, delimiter, n, "split")

    def rsplit(self, delimiter=None, n=-1):
        """reference nvstrings.py:1099"""
        return self._split(lib().custr_rsplit, delimiter, n, "rsplit")

    def _split_record(self, fn, delimiter, n, what):
        rows = self.size()
        row_off = np.zeros(rows + 1, np.int32)
        tok = C.c_void_p()
        total = check_rc(fn(self.m_cptr, _enc(delimiter), int(n), C.byref(tok), as_ptr(row_off), 0), what)
        if total is not None and total == -1:
            raise ValueError(what + ": " + _lib.last_error())
        tokens = nvstrings(check_handle(tok.value, what))
        valid = self._valid_mask()
        res = []
        for i in range(rows):
            if not valid[i]:
                res.append(None)
            else:
                h = lib().custr_slice_rows(tokens.m_cptr, int(row_off[i]), int(row_off[i + 1]))
                res.append(nvstrings(check_handle(h, what)))
        return res

    def split_record(self, delimiter=None, n=-1):
        """Row-major split: one nvstrings per input row (None for null rows), all views over ONE flat token column
        (the reference allocates N objects, split.cu:171-190).  reference nvstrings.py:936"""
        return self._split_record(lib().custr_split_record, delimiter, n, "split_record")

    def rsplit_record(self, delimiter=None, n=-1):
        """reference nvstrings.py:969"""
        return self._split_record(lib().custr_rsplit_record, delimiter, n, "rsplit_record")

    def _partition(self, delimiter, right, what):
        if delimiter is None or delimiter == "" or delimiter == b"":
            return []  # reference returns without results (split.cu:1167-1171)
        h = lib().custr_partition(self.m_cptr, _enc(delimiter), 1 if right else 0)
        flat = nvstrings(check_handle(h, what))
        res = []
        for i in range(self.size()):
            res.append(nvstrings(check_handle(lib().custr_slice_rows(flat.m_cptr, 3 * i, 3 * i + 3), what)))
        return res

    def partition(self, delimiter=" "):
        """One nvstrings of 3 strings per row: [left, delimiter, right] around the first delimiter ([row, '', ''] when it
        does not occur, three nulls for a null row).  reference nvstrings.py:1127 -> split.cu:1165-1262"""
        return self._partition(delimiter, False, "partition")

    def rpartition(self, delimiter=" "):
        """Same around the LAST delimiter (['', '', row] when it does not occur).  reference nvstrings.py:1163 -> split.cu:1268-1372"""
        return self._partition(delimiter, True, "rpartition")

    def partition_flat(self, delimiter=" ", right=False):
        """The 3n-row column partition()/rpartition() are views of."""
        h = lib().custr_partition(self.m_cptr, _enc(delimiter), 1 if right else 0)
        return nvstrings(check_handle(h, "partition"))

    def split_record_flat(self, delimiter=None, n=-1, right=False):
        """(tokens nvstrings, row_offsets int32[n+1]) — the flat form split_record is built on."""
        rows = self.size()
        row_off = np.zeros(rows + 1, np.int32)
        tok = C.c_void_p()
        fn = lib().custr_rsplit_record if right else lib().custr_split_record
        check_rc(fn(self.m_cptr, _enc(delimiter), int(n), C.byref(tok), as_ptr(row_off), 0), "split_record")
        return nvstrings(check_handle(tok.value, "split_record")), row_off

    def __getattr__(self, name):
        if name.startswith("_") or name in ("m_cptr",):
            raise AttributeError(name)
        raise NotImplementedError(
            "nvstrings.%s is outside the hot path implemented by custrings_b200 (SURVEY.md section 8)" % name)

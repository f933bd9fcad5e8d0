"""ctypes front-end to oracle/_ref/libref_oracle.so — the UNMODIFIED reference (rapidsai/custrings
cpp/src) compiled for the host CPU by oracle/Makefile, wrapped by oracle/ref_harness.cpp.

TEST INFRASTRUCTURE ONLY.  Importers allowed: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline /
--impl reference legs.  The product (custrings_b200/) never imports this module.

Columns are exchanged as Arrow-style triples: chars uint8[total], offsets int32[n+1], validity uint8[(n+7)//8]
(LSB-first, bit=1 => valid) — the reference's own import/export format
(cpp/include/NVStrings.h:116 create_from_offsets, :207 create_offsets).
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libref_oracle.so")
_lib = None


def available():
    return os.path.exists(_SO)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libref_oracle.so missing: run `make -C oracle` (needs /root/reference)")
        L = C.CDLL(_SO)
        vp, ci, cp = C.c_void_p, C.c_int, C.c_char_p
        L.ref_last_error.restype = cp
        for name in ("ref_create", "ref_create_from_array", "ref_replace_re", "ref_replace_re_multi", "ref_replace", "ref_replace_with_backrefs",
                     "ref_replace_multi", "ref_tokenize", "ref_tokenize_multi", "ref_cat_create", "ref_cat_create_multi",
                     "ref_cat_keys", "ref_cat_to_strings", "ref_cat_merge", "ref_cat_from_categories", "ref_cat_keys_op", "ref_cat_gather",
                     "ref_cat_gather_strings", "ref_case", "ref_strip", "ref_slice"):
            getattr(L, name).restype = vp
        L.ref_create.argtypes = [vp, ci, vp, vp, ci]
        L.ref_destroy.argtypes = [vp]
        L.ref_size.argtypes = [vp]; L.ref_size.restype = C.c_uint
        L.ref_total_bytes.argtypes = [vp]; L.ref_total_bytes.restype = C.c_long
        L.ref_export.argtypes = [vp, vp, vp, vp]
        L.ref_len.argtypes = [vp, vp]
        L.ref_hash.argtypes = [vp, vp]
        for name in ("ref_contains_re", "ref_match", "ref_count_re", "ref_contains", "ref_startswith", "ref_endswith"):
            getattr(L, name).argtypes = [vp, cp, vp]
        L.ref_replace_re.argtypes = [vp, cp, cp, ci]
        L.ref_replace_with_backrefs.argtypes = [vp, cp, cp]
        L.ref_replace_re_multi.argtypes = [vp, vp, ci, vp]
        L.ref_replace.argtypes = [vp, cp, cp, ci]
        L.ref_replace_multi.argtypes = [vp, vp, vp]
        L.ref_find.argtypes = [vp, cp, ci, ci, vp]
        L.ref_rfind.argtypes = [vp, cp, ci, ci, vp]
        L.ref_find_from.argtypes = [vp, cp, vp, vp, vp]
        L.ref_match_strings.argtypes = [vp, vp, vp]
        L.ref_find_multiple.argtypes = [vp, vp, vp]
        L.ref_split.argtypes = [vp, cp, ci, ci, vp, ci]
        L.ref_split_record.argtypes = [vp, cp, ci, ci, vp]
        L.ref_partition.argtypes = [vp, cp, ci, vp]
        L.ref_regex_columns.argtypes = [vp, cp, ci, vp, ci]
        L.ref_regex_records.argtypes = [vp, cp, ci, vp]
        L.ref_regex_dump.argtypes = [cp, vp, ci]
        L.ref_tokenize.argtypes = [vp, cp]
        L.ref_tokenize_multi.argtypes = [vp, vp]
        L.ref_token_count.argtypes = [vp, cp, vp]
        L.ref_cat_create.argtypes = [vp]
        L.ref_cat_create_multi.argtypes = [vp, ci]
        L.ref_cat_destroy.argtypes = [vp]
        L.ref_cat_size.argtypes = [vp]; L.ref_cat_size.restype = C.c_uint
        L.ref_cat_keys_size.argtypes = [vp]; L.ref_cat_keys_size.restype = C.c_uint
        L.ref_cat_keys.argtypes = [vp]
        L.ref_cat_values.argtypes = [vp, vp]
        L.ref_cat_to_strings.argtypes = [vp]
        L.ref_cat_merge.argtypes = [vp, vp, ci]
        L.ref_cat_from_categories.argtypes = [vp, ci]
        L.ref_cat_keys_op.argtypes = [vp, vp, ci]
        L.ref_is_class.argtypes = [vp, ci, vp]
        L.ref_case.argtypes = [vp, ci]
        L.ref_strip.argtypes = [vp, cp, ci]
        L.ref_slice.argtypes = [vp, ci, ci, ci]
        L.ref_cat_gather.argtypes = [vp, vp, C.c_uint, ci]
        L.ref_cat_gather_strings.argtypes = [vp, vp, C.c_uint]
        _lib = L
    return _lib


def _b(s):
    if s is None:
        return None
    return s.encode("utf-8") if isinstance(s, str) else bytes(s)


def pack(strings):
    """list of str/bytes/None -> (chars, offsets, validity, nulls)"""
    n = len(strings)
    enc = [None if s is None else _b(s) for s in strings]
    offsets = np.zeros(n + 1, dtype=np.int32)
    lens = np.fromiter((0 if e is None else len(e) for e in enc), dtype=np.int64, count=n)
    np.cumsum(lens, out=offsets[1:])
    chars = np.frombuffer(b"".join(e for e in enc if e is not None), dtype=np.uint8).copy()
    valid = np.fromiter((e is not None for e in enc), dtype=bool, count=n)
    validity = np.packbits(valid, bitorder="little") if n else np.zeros(0, np.uint8)
    return chars, offsets, validity, int(n - valid.sum())


def unpack(chars, offsets, validity):
    """-> list of bytes/None"""
    n = len(offsets) - 1
    if validity is None:
        valid = np.ones(n, bool)
    else:
        valid = np.unpackbits(np.asarray(validity, np.uint8), bitorder="little")[:n].astype(bool)
    raw = chars.tobytes()
    return [raw[offsets[i]:offsets[i + 1]] if valid[i] else None for i in range(n)]


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class RefStrings:
    """Owning handle on a reference NVStrings instance (host build)."""

    def __init__(self, handle):
        if not handle:
            raise ValueError("reference error: " + lib().ref_last_error().decode())
        self.h = C.c_void_p(handle)

    def __del__(self):
        try:
            if self.h:
                lib().ref_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---- construction / export
    @staticmethod
    def from_arrays(chars, offsets, validity=None, nulls=0):
        chars = np.ascontiguousarray(chars, np.uint8)
        if chars.size == 0:
            chars = np.zeros(1, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.int32)
        if validity is not None:
            validity = np.ascontiguousarray(validity, np.uint8)
            if nulls == 0:
                n = len(offsets) - 1
                nulls = int(n - np.unpackbits(validity, bitorder="little")[:n].sum())
        return RefStrings(lib().ref_create(_ptr(chars), len(offsets) - 1, _ptr(offsets), _ptr(validity) if nulls else None, nulls))

    @staticmethod
    def from_list(strings):
        if len(strings) == 0:
            return RefStrings.from_arrays(np.zeros(0, np.uint8), np.zeros(1, np.int32))
        chars, offsets, validity, nulls = pack(strings)
        return RefStrings.from_arrays(chars, offsets, validity, nulls)

    def size(self):
        return int(lib().ref_size(self.h))

    def to_arrays(self):
        n = self.size()
        total = int(lib().ref_total_bytes(self.h))
        chars = np.zeros(max(total, 1), np.uint8)
        offsets = np.zeros(n + 1, np.int32)
        validity = np.zeros((n + 7) // 8, np.uint8)
        if n:
            rc = lib().ref_export(self.h, _ptr(chars), _ptr(offsets), _ptr(validity))
            if rc == -100:
                raise ValueError(lib().ref_last_error().decode())
        return chars[:total], offsets, validity

    def to_list(self):
        return unpack(*self.to_arrays())

    # ---- per-row scalar results
    def _rows(self, fn, dtype, *args):
        out = np.zeros(max(self.size(), 1), dtype)
        rc = fn(self.h, *args, _ptr(out))
        if rc == -100:
            raise ValueError(lib().ref_last_error().decode())
        return out[: self.size()], rc

    def contains_re(self, pat): return self._rows(lib().ref_contains_re, np.bool_, _b(pat))
    def match(self, pat): return self._rows(lib().ref_match, np.bool_, _b(pat))
    def count_re(self, pat): return self._rows(lib().ref_count_re, np.int32, _b(pat))
    def contains(self, s): return self._rows(lib().ref_contains, np.bool_, _b(s))
    def startswith(self, s): return self._rows(lib().ref_startswith, np.bool_, _b(s))
    def endswith(self, s): return self._rows(lib().ref_endswith, np.bool_, _b(s))
    def find(self, s, start=0, end=-1): return self._rows(lib().ref_find, np.int32, _b(s), start, end)
    def rfind(self, s, start=0, end=-1): return self._rows(lib().ref_rfind, np.int32, _b(s), start, end)
    def len(self): return self._rows(lib().ref_len, np.int32)
    def hash(self): return self._rows(lib().ref_hash, np.uint32)
    def token_count(self, delim=None): return self._rows(lib().ref_token_count, np.uint32, _b(delim))

    def find_from(self, s, starts=None, ends=None):
        hs = None if starts is None else np.ascontiguousarray(starts, np.int32)
        he = None if ends is None else np.ascontiguousarray(ends, np.int32)
        out = np.zeros(max(self.size(), 1), np.int32)
        rc = lib().ref_find_from(self.h, _b(s), None if hs is None else _ptr(hs), None if he is None else _ptr(he), _ptr(out))
        return out[: self.size()], rc

    def match_strings(self, other):
        out = np.zeros(max(self.size(), 1), np.bool_)
        rc = lib().ref_match_strings(self.h, other.h, _ptr(out))
        return out[: self.size()], rc

    def find_multiple(self, targets):
        out = np.zeros(max(self.size() * targets.size(), 1), np.int32)
        rc = lib().ref_find_multiple(self.h, targets.h, _ptr(out))
        return out[: self.size() * targets.size()].reshape(self.size(), targets.size()), rc

    # ---- column results
    def replace_re(self, pat, repl, maxrepl=-1):
        return RefStrings(lib().ref_replace_re(self.h, _b(pat), _b(repl), maxrepl))

    def replace_with_backrefs(self, pat, repl):
        h = lib().ref_replace_with_backrefs(self.h, _b(pat), _b(repl))
        if not h:
            raise ValueError(lib().ref_last_error().decode())
        return RefStrings(h)

    def replace_re_multi(self, pats, repls):
        arr = (C.c_char_p * len(pats))(*[_b(p) for p in pats])
        return RefStrings(lib().ref_replace_re_multi(self.h, arr, len(pats), repls.h))

    def replace(self, s, repl, maxrepl=-1):
        return RefStrings(lib().ref_replace(self.h, _b(s), _b(repl), maxrepl))

    def replace_multi(self, tgts, repls):
        return RefStrings(lib().ref_replace_multi(self.h, tgts.h, repls.h))

    def split(self, delim=None, maxsplit=-1, right=False):
        cap = 4096
        out = (C.c_void_p * cap)()
        n = lib().ref_split(self.h, _b(delim), maxsplit, int(right), out, cap)
        if n == -100:
            raise ValueError(lib().ref_last_error().decode())
        return [RefStrings(out[i]) for i in range(min(n, cap))]

    def split_record(self, delim=None, maxsplit=-1, right=False):
        n = self.size()
        out = (C.c_void_p * max(n, 1))()
        total = lib().ref_split_record(self.h, _b(delim), maxsplit, int(right), out)
        if total == -100:
            raise ValueError(lib().ref_last_error().decode())
        return [RefStrings(out[i]) if out[i] else None for i in range(n)], total

    def partition(self, delim, right=False):
        n = self.size()
        out = (C.c_void_p * max(n, 1))()
        rc = lib().ref_partition(self.h, _b(delim), int(right), out)
        return [RefStrings(out[i]) if out[i] else None for i in range(n)], rc

    def _regex_columns(self, pat, kind):
        cap = 256
        out = (C.c_void_p * cap)()
        n = lib().ref_regex_columns(self.h, _b(pat), kind, out, cap)
        if n == -100:
            raise ValueError(lib().ref_last_error().decode())
        return [RefStrings(out[i]) for i in range(min(n, cap))]

    def findall(self, pat): return self._regex_columns(pat, 0)
    def extract(self, pat): return self._regex_columns(pat, 1)

    def _regex_records(self, pat, kind):
        n = self.size()
        out = (C.c_void_p * max(n, 1))()
        rc = lib().ref_regex_records(self.h, _b(pat), kind, out)
        if rc == -100:
            raise ValueError(lib().ref_last_error().decode())
        return [RefStrings(out[i]) if out[i] else None for i in range(n)]

    def findall_record(self, pat): return self._regex_records(pat, 0)
    def extract_record(self, pat): return self._regex_records(pat, 1)

    def tokenize(self, delim=None):
        return RefStrings(lib().ref_tokenize(self.h, _b(delim)))

    def is_class(self, kind): return self._rows(lib().ref_is_class, np.bool_, kind)
    def case(self, upper): return RefStrings(lib().ref_case(self.h, 1 if upper else 0))
    def strip(self, chars=None, side=0): return RefStrings(lib().ref_strip(self.h, _b(chars), side))
    def slice(self, start, stop=-1, step=1): return RefStrings(lib().ref_slice(self.h, start, stop, step))

    def tokenize_multi(self, delims):
        return RefStrings(lib().ref_tokenize_multi(self.h, delims.h))


class RefCategory:
    def __init__(self, strs, handle=None):
        if handle is not None:
            self.h = C.c_void_p(handle)
            return
        if isinstance(strs, (list, tuple)):
            arr = (C.c_void_p * len(strs))(*[s.h for s in strs])
            h = lib().ref_cat_create_multi(arr, len(strs))
        else:
            h = lib().ref_cat_create(strs.h)
        if not h:
            raise ValueError(lib().ref_last_error().decode())
        self.h = C.c_void_p(h)

    def __del__(self):
        try:
            if self.h:
                lib().ref_cat_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def size(self): return int(lib().ref_cat_size(self.h))
    def keys_size(self): return int(lib().ref_cat_keys_size(self.h))
    def keys(self): return RefStrings(lib().ref_cat_keys(self.h))
    def to_strings(self): return RefStrings(lib().ref_cat_to_strings(self.h))
    def merge_category(self, other): return RefCategory(None, lib().ref_cat_merge(self.h, other.h, 0))
    def merge_and_remap(self, other): return RefCategory(None, lib().ref_cat_merge(self.h, other.h, 1))

    def _cat(self, h):
        if not h:
            raise ValueError(lib().ref_last_error().decode())
        return RefCategory(None, h)

    def add_keys(self, strs): return self._cat(lib().ref_cat_keys_op(self.h, strs.h, 0))
    def remove_keys(self, strs): return self._cat(lib().ref_cat_keys_op(self.h, strs.h, 1))
    def set_keys(self, strs): return self._cat(lib().ref_cat_keys_op(self.h, strs.h, 2))
    def remove_unused_keys(self): return self._cat(lib().ref_cat_keys_op(self.h, None, 3))

    def gather(self, pos, remap=False):
        a = np.ascontiguousarray(pos, np.int32)
        return self._cat(lib().ref_cat_gather(self.h, _ptr(a), len(a), 1 if remap else 0))

    def gather_strings(self, pos):
        a = np.ascontiguousarray(pos, np.int32)
        h = lib().ref_cat_gather_strings(self.h, _ptr(a), len(a))
        if not h:
            raise ValueError(lib().ref_last_error().decode())
        return RefStrings(h)

    @staticmethod
    def from_categories(cats):
        arr = (C.c_void_p * len(cats))(*[c.h for c in cats])
        return RefCategory(None, lib().ref_cat_from_categories(arr, len(cats)))

    def values(self):
        out = np.zeros(max(self.size(), 1), np.int32)
        lib().ref_cat_values(self.h, _ptr(out))
        return out[: self.size()]

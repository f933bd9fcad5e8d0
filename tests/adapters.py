"""Two adapters with the same surface so that one KAT / parity table can be replayed against (a) the oracle
(reference compiled for the CPU) and (b) the product through the C-ABI on the GPU."""
import numpy as np


def _b2s(x):
    return None if x is None else x.decode("utf-8")


class OracleAPI:
    def __init__(self, ref):
        self.ref = ref

    def column(self, strs):
        return self.ref.RefStrings.from_list(strs)

    def strings(self, col):
        return [_b2s(x) for x in col.to_list()]

    def contains(self, c, s): return c.contains(s)[0].tolist()
    def contains_re(self, c, p): return c.contains_re(p)[0].tolist()
    def match(self, c, p): return c.match(p)[0].tolist()
    def count_re(self, c, p): return c.count_re(p)[0].tolist()
    def replace(self, c, s, r, n=-1): return self.strings(c.replace(s, r, n))
    def replace_re(self, c, p, r, n=-1): return self.strings(c.replace_re(p, r, n))
    def replace_multi(self, c, t, r): return self.strings(c.replace_multi(self.column(t), self.column(r)))
    def replace_re_multi(self, c, p, r): return self.strings(c.replace_re_multi(p, self.column(r)))
    def find(self, c, s, a, b): return c.find(s, a, b)[0].tolist()
    def rfind(self, c, s, a, b): return c.rfind(s, a, b)[0].tolist()
    def find_multiple(self, c, t): return c.find_multiple(self.column(t))[0].tolist()
    def startswith(self, c, s): return c.startswith(s)[0].tolist()
    def endswith(self, c, s): return c.endswith(s)[0].tolist()
    def split(self, c, d, n): return [self.strings(x) for x in c.split(d, n)]
    def split_record(self, c, d, n): return [None if r is None else self.strings(r) for r in c.split_record(d, n)[0]]
    def rsplit(self, c, d, n): return [self.strings(x) for x in c.split(d, n, right=True)]
    def rsplit_record(self, c, d, n): return [None if r is None else self.strings(r) for r in c.split_record(d, n, right=True)[0]]
    def partition(self, c, d): return [self.strings(r) for r in c.partition(d)[0]]
    def rpartition(self, c, d): return [self.strings(r) for r in c.partition(d, right=True)[0]]
    def replace_with_backrefs(self, c, p, r): return self.strings(c.replace_with_backrefs(p, r))
    def tokenize(self, c, d): return self.strings(c.tokenize(d))
    def token_count(self, c, d): return c.token_count(d)[0].tolist()
    def hash(self, c): return c.hash()[0].tolist()

    def category(self, c):
        cat = self.ref.RefCategory(c)
        return (self.strings(cat.keys()), cat.values().tolist())


class ProductAPI:
    """custrings_b200 through ctypes -> libcustr.so (GPU)."""

    def __init__(self):
        from custrings_b200 import nvstrings, nvcategory, nvtext
        self.nvs, self.nvc, self.nvt = nvstrings, nvcategory, nvtext

    def column(self, strs):
        return self.nvs.to_device(strs)

    def strings(self, col):
        return col.to_host()

    @staticmethod
    def _f(lst, fill):
        return [fill if x is None else x for x in lst]

    def contains(self, c, s): return self._f(c.contains(s, regex=False), False)
    def contains_re(self, c, p): return self._f(c.contains(p), False)
    def match(self, c, p): return self._f(c.match(p), False)
    def count_re(self, c, p): return self._f(c.count(p), 0)
    def replace(self, c, s, r, n=-1): return c.replace(s, r, n, regex=False).to_host()
    def replace_re(self, c, p, r, n=-1): return c.replace(p, r, n).to_host()
    def replace_multi(self, c, t, r): return c.replace_multi(list(t), self.column(r), regex=False).to_host()
    def replace_re_multi(self, c, p, r): return c.replace_multi(list(p), self.column(r), regex=True).to_host()
    def find(self, c, s, a, b): return self._f(c.find(s, a, None if b == -1 else b), -2)
    def rfind(self, c, s, a, b): return self._f(c.rfind(s, a, None if b == -1 else b), -2)
    def find_multiple(self, c, t): return c.find_multiple(self.column(t))
    def startswith(self, c, s): return self._f(c.startswith(s), False)
    def endswith(self, c, s): return self._f(c.endswith(s), False)
    def split(self, c, d, n): return [x.to_host() for x in c.split(d, n)]
    def split_record(self, c, d, n): return [None if r is None else r.to_host() for r in c.split_record(d, n)]
    def rsplit(self, c, d, n): return [x.to_host() for x in c.rsplit(d, n)]
    def rsplit_record(self, c, d, n): return [None if r is None else r.to_host() for r in c.rsplit_record(d, n)]
    def partition(self, c, d): return [r.to_host() for r in c.partition(d)]
    def rpartition(self, c, d): return [r.to_host() for r in c.rpartition(d)]
    def replace_with_backrefs(self, c, p, r): return c.replace_with_backrefs(p, r).to_host()
    def tokenize(self, c, d): return self.nvt.tokenize(c, d).to_host()
    def token_count(self, c, d): return self.nvt.token_count(c, d)
    def hash(self, c): return self._f(c.hash(), 0)

    def category(self, c):
        cat = self.nvc.from_strings(c)
        return (cat.keys().to_host(), cat.values())

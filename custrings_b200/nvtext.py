"""nvtext — host-side mirror of the reference shim python/nvtext.py for tokenize / token_count
(nvtext.py:7-44, :76) over libcustr.so's C-ABI."""
import numpy as np

from ._lib import as_ptr, check_handle, check_rc, lib
from . import nvstrings as _nvs


def _enc(s):
    return None if s is None else (s.encode("utf-8") if isinstance(s, str) else bytes(s))


def tokenize(strs, delimiter=None):
    """All tokens of all strings as one nvstrings, in row order.  delimiter None = whitespace, else every character
    of `delimiter` separates tokens.  reference nvtext.py:7 -> tokens.cu:123-155"""
    if strs is None:
        raise ValueError("tokenize: strs is None")
    if delimiter is not None and not isinstance(delimiter, (str, bytes)):
        raise NotImplementedError("tokenize with a list of multi-character delimiters is outside the hot path")
    h = lib().custr_tokenize(strs.m_cptr, _enc(delimiter))
    return _nvs.nvstrings(check_handle(h, "tokenize"))


def token_count(strs, delimiter=None, devptr=0):
    """Number of tokens per string.  reference nvtext.py:76 -> tokens.cu:337-361"""
    n = strs.size()
    if devptr:
        check_rc(lib().custr_token_count(strs.m_cptr, _enc(delimiter), as_ptr(devptr), 1), "token_count")
        return devptr
    out = np.zeros(max(n, 1), np.uint32)
    check_rc(lib().custr_token_count(strs.m_cptr, _enc(delimiter), as_ptr(out), 0), "token_count")
    return out[:n].tolist()

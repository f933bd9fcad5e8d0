"""ctypes access to the host simulation harness (tests/sim) — test infrastructure."""
import ctypes as C
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
_lib = None


def lib():
    global _lib
    if _lib is None:
        from tests.sim.build_sim import build
        _lib = C.CDLL(build())
        _lib.sim_replace_re.restype = C.c_long
        _lib.sim_replace_re_multi.restype = C.c_long
        _lib.sim_chain_replace.restype = C.c_long
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _cols(chars, offsets, validity):
    chars = np.ascontiguousarray(chars, np.uint8)
    if chars.size == 0:
        chars = np.zeros(1, np.uint8)
    offsets = np.ascontiguousarray(offsets, np.int32)
    validity = None if validity is None else np.ascontiguousarray(validity, np.uint8)
    return chars, offsets, validity


def describe(pattern):
    buf = C.create_string_buffer(1 << 16)
    lib().sim_describe(pattern.encode() if isinstance(pattern, str) else pattern, buf, len(buf))
    return buf.value.decode(errors="replace")


def bool_search(chars, offsets, validity, pattern, anchored):
    chars, offsets, validity = _cols(chars, offsets, validity)
    n = len(offsets) - 1
    out = np.zeros(max(n, 1), np.uint8)
    rc = lib().sim_bool(_p(chars), _p(offsets), _p(validity), n, pattern.encode() if isinstance(pattern, str) else pattern, int(anchored), _p(out))
    return out[:n].astype(bool), rc


def bits_bool(chars, offsets, validity, pattern, anchored):
    """bitstream tier simulated on the host; returns (None, -1) when the pattern is not eligible"""
    chars, offsets, validity = _cols(chars, offsets, validity)
    n = len(offsets) - 1
    out = np.zeros(max(n, 1), np.uint8)
    rc = lib().sim_bits_bool(_p(chars), _p(offsets), _p(validity), n, pattern.encode() if isinstance(pattern, str) else pattern, int(anchored), _p(out))
    if rc < 0:
        return None, -1
    return out[:n].astype(bool), rc


def chain_bool(chars, offsets, validity, pattern, anchored):
    """the chain model (steps with loop / opt / exit flags) executed on the host; (None, -1) when the pattern is not a chain"""
    chars, offsets, validity = _cols(chars, offsets, validity)
    n = len(offsets) - 1
    out = np.zeros(max(n, 1), np.uint8)
    rc = lib().sim_chain_bool(_p(chars), _p(offsets), _p(validity), n, pattern.encode() if isinstance(pattern, str) else pattern, int(anchored), _p(out))
    if rc < 0:
        return None, -1
    return out[:n].astype(bool), rc


def chain_count(chars, offsets, validity, pattern):
    """span fast path; (None, -1) when the pattern is not a last-loop chain"""
    chars, offsets, validity = _cols(chars, offsets, validity)
    n = len(offsets) - 1
    out = np.zeros(max(n, 1), np.int32)
    rc = lib().sim_chain_count(_p(chars), _p(offsets), _p(validity), n, pattern.encode() if isinstance(pattern, str) else pattern, _p(out))
    return (None, -1) if rc < 0 else (out[:n], rc)


def chain_replace(chars, offsets, validity, pattern, repl, maxrepl=-1):
    chars, offsets, validity = _cols(chars, offsets, validity)
    n = len(offsets) - 1
    pat = pattern.encode() if isinstance(pattern, str) else pattern
    rp = repl.encode() if isinstance(repl, str) else repl
    ooff = np.zeros(n + 1, np.int32)
    total = lib().sim_chain_replace(_p(chars), _p(offsets), _p(validity), n, pat, rp, maxrepl, _p(ooff), None)
    if total < 0:
        return None
    ochars = np.zeros(max(total, 1), np.uint8)
    lib().sim_chain_replace(_p(chars), _p(offsets), _p(validity), n, pat, rp, maxrepl, _p(ooff), _p(ochars))
    return ochars[:total], ooff


def count(chars, offsets, validity, pattern):
    chars, offsets, validity = _cols(chars, offsets, validity)
    n = len(offsets) - 1
    out = np.zeros(max(n, 1), np.int32)
    rc = lib().sim_count(_p(chars), _p(offsets), _p(validity), n, pattern.encode() if isinstance(pattern, str) else pattern, _p(out))
    return out[:n], rc


def replace_re(chars, offsets, validity, pattern, repl, maxrepl=-1):
    chars, offsets, validity = _cols(chars, offsets, validity)
    n = len(offsets) - 1
    pat = pattern.encode() if isinstance(pattern, str) else pattern
    rp = repl.encode() if isinstance(repl, str) else repl
    ooff = np.zeros(n + 1, np.int32)
    total = lib().sim_replace_re(_p(chars), _p(offsets), _p(validity), n, pat, rp, maxrepl, _p(ooff), None)
    ochars = np.zeros(max(total, 1), np.uint8)
    lib().sim_replace_re(_p(chars), _p(offsets), _p(validity), n, pat, rp, maxrepl, _p(ooff), _p(ochars))
    return ochars[:total], ooff


def replace_re_multi(chars, offsets, validity, patterns, rchars, roffsets, rvalidity):
    chars, offsets, validity = _cols(chars, offsets, validity)
    rchars, roffsets, rvalidity = _cols(rchars, roffsets, rvalidity)
    n = len(offsets) - 1
    arr = (C.c_char_p * len(patterns))(*[p.encode() if isinstance(p, str) else p for p in patterns])
    ooff = np.zeros(n + 1, np.int32)
    args = (_p(chars), _p(offsets), _p(validity), n, arr, len(patterns), _p(rchars), _p(roffsets), _p(rvalidity), len(roffsets) - 1)
    total = lib().sim_replace_re_multi(*args, _p(ooff), None)
    ochars = np.zeros(max(total, 1), np.uint8)
    lib().sim_replace_re_multi(*args, _p(ooff), _p(ochars))
    return ochars[:total], ooff


def available():
    try:
        lib()
        return True
    except Exception:
        return False


def program_dump(pattern):
    """this repo's compiled program in the layout of oracle.restate.reference_program"""
    out = np.zeros(1 << 16, np.int32)
    n = lib().sim_program_dump(pattern.encode() if isinstance(pattern, str) else pattern, _p(out), len(out))
    w = out[:n].tolist()
    ninsts, start, ngroups, nstarts, nclasses = w[:5]
    p = 5
    insts = [tuple(w[p + 3 * i: p + 3 * i + 3]) for i in range(ninsts)]
    p += 3 * ninsts
    starts = w[p:p + nstarts]
    p += nstarts
    classes = []
    for _ in range(nclasses):
        builtins, cnt = w[p], w[p + 1]
        classes.append((builtins, [x & 0xFFFFFFFF for x in w[p + 2:p + 2 + cnt]]))
        p += 2 + cnt
    return {"insts": insts, "start": start, "groups": ngroups, "starts": starts, "classes": classes}

"""nvcategory — host-side mirror of the reference shim python/nvcategory.py for the key-build path
(from_strings :78, keys :242, values :364, keys_size, size), over libcustr.so's C-ABI.

Multi-GPU: from_strings_sharded() builds the local dictionary of a row shard, all-gathers the
distinct keys of every rank (ncclAllGather over NVLink inside libcustr.so's host C++; torch.distributed tensor collectives
with gloo in the CPU tests of the host logic) and remaps the local values onto the global sorted key set.
"""
import ctypes as C

import numpy as np

from ._lib import as_ptr, check_handle, check_rc, lib
from . import nvstrings as _nvs


def from_strings(*args):
    """Create an nvcategory from one or more nvstrings.  reference nvcategory.py:78 -> NVCategory.cu:327-356"""
    cols = []
    for a in args:
        cols.extend(a if isinstance(a, (list, tuple)) else [a])
    for c in cols:
        if type(c).__name__ != "nvstrings":  # reference checks the type NAME, pycategory.cpp:42-71
            raise ValueError("from_strings: arguments must be nvstrings")
    arr = (C.c_void_p * len(cols))(*[c.m_cptr for c in cols])
    h = lib().custr_category_create(arr, len(cols))
    return nvcategory(check_handle(h, "from_strings"))


def to_device(strs):
    """reference nvcategory.py:7"""
    return from_strings(_nvs.to_device(strs))


def from_offsets(sbuf, obuf, scount, nbuf=None, ncount=0, bdevmem=False):
    """reference nvcategory.py:37 -> NVCategory.cu:359-370"""
    return from_strings(_nvs.from_offsets(sbuf, obuf, scount, nbuf, ncount, bdevmem))


def exchange_key_arrays(chars, offsets, valid, group=None, device=None):
    """All-gather one (chars uint8[], offsets int32[k+1], valid bool[k]) key column per rank with tensor collectives
    (NCCL over NVLink on GPUs, gloo in the CPU tests): sizes first, then max-padded payloads.  Pure host/torch logic,
    no custr calls.  Returns the rank-ordered concatenation (chars, offsets, valid)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return chars, offsets, valid
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    n = len(offsets) - 1
    meta = torch.tensor([n, int(offsets[-1]) - int(offsets[0])], dtype=torch.int64, device=device)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    metas = [m.cpu().tolist() for m in metas]
    max_n = max(m[0] for m in metas)
    max_b = max(m[1] for m in metas)
    lens = np.zeros(max_n + 1, np.int32)
    lens[:n] = np.diff(offsets)
    vflags = np.zeros(max_n + 1, np.int32)
    vflags[:n] = valid
    payload = np.zeros(max_b + 1, np.uint8)
    payload[: metas[dist.get_rank(group)][1]] = chars[int(offsets[0]): int(offsets[-1])]
    t_len, t_val, t_chr = (torch.from_numpy(x).to(device) for x in (lens, vflags, payload))
    g_len = [torch.zeros_like(t_len) for _ in range(world)]
    g_val = [torch.zeros_like(t_val) for _ in range(world)]
    g_chr = [torch.zeros_like(t_chr) for _ in range(world)]
    dist.all_gather(g_len, t_len, group=group)
    dist.all_gather(g_val, t_val, group=group)
    dist.all_gather(g_chr, t_chr, group=group)
    all_lens = np.concatenate([g_len[r][: metas[r][0]].cpu().numpy() for r in range(world)])
    all_valid = np.concatenate([g_val[r][: metas[r][0]].cpu().numpy().astype(bool) for r in range(world)])
    all_chars = np.concatenate([g_chr[r][: metas[r][1]].cpu().numpy() for r in range(world)])
    offs = np.zeros(len(all_lens) + 1, np.int32)
    np.cumsum(all_lens, out=offs[1:])
    return all_chars, offs, all_valid


def gather_keys_device(keys, group=None):
    """All-gather an nvstrings of keys across ranks; returns an nvstrings holding every rank's keys in rank order."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return keys
    chars, offsets, validity = keys.to_arrays()
    n = len(offsets) - 1
    valid = np.unpackbits(validity, bitorder="little")[:n].astype(bool) if n else np.zeros(0, bool)
    chars, offs, valid = exchange_key_arrays(chars, offsets, valid, group)
    nulls = int((~valid).sum())
    return _nvs.from_offsets(chars if chars.size else np.zeros(1, np.uint8), offs, len(offs) - 1, np.packbits(valid, bitorder="little"), nulls)


_COMMS = {}  # (id(group) or None) -> custr_comm handle of this process


def _comm(group=None):
    """NCCL communicator of libcustr.so for `group` (created once per process and group): rank 0 draws the unique id,
    torch.distributed only carries its 128 bytes to the other ranks."""
    import torch
    import torch.distributed as dist
    key = id(group) if group is not None else None
    if key not in _COMMS:
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        buf = (C.c_ubyte * 128)()
        if rank == 0:
            check_rc(lib().custr_comm_unique_id(buf), "comm_unique_id")
        t = torch.tensor(list(buf), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        ident = (C.c_ubyte * 128)(*t.cpu().tolist())
        _COMMS[key] = check_handle(lib().custr_comm_create(rank, world, ident), "comm_create")
    return _COMMS[key]


def from_strings_sharded(strs, group=None, timings=False):
    """Row-sharded dictionary build: `strs` is THIS rank's contiguous row range.  Every rank returns a category whose
    keys() is the global sorted key set and whose values() index into it.  On GPUs (NCCL backend) the whole exchange runs in
    host C++ inside libcustr.so (custr_category_create_sharded: ncclAllGather of the ranks' distinct keys); with the gloo
    backend (CPU tests of the host logic) the keys travel through torch collectives (exchange_key_arrays).
    timings=True: returns (category, {build_ms, exchange_ms, remap_ms, how})."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1 and dist.get_backend(group) == "nccl":
        t = (C.c_float * 3)()
        h = lib().custr_category_create_sharded(_comm(group), strs.m_cptr, t)
        cat = nvcategory(check_handle(h, "from_strings_sharded"))
        info = {"build_ms": t[0], "exchange_ms": t[1], "remap_ms": t[2], "how": "ncclAllGather x3 inside libcustr.so (sizes, key lengths, key bytes)"}
        return (cat, info) if timings else cat
    import time
    t0 = time.perf_counter()
    local = from_strings(strs)
    t1 = time.perf_counter()
    all_keys = gather_keys_device(local.keys(), group)
    t2 = time.perf_counter()
    h = lib().custr_category_remap_to_union(local.m_cptr, all_keys.m_cptr)
    cat = nvcategory(check_handle(h, "from_strings_sharded"))
    info = {"build_ms": 1e3 * (t1 - t0), "exchange_ms": 1e3 * (t2 - t1), "remap_ms": 1e3 * (time.perf_counter() - t2), "how": "torch.distributed tensor collectives"}
    return (cat, info) if timings else cat


def from_categories(cats):
    """NVCategory::create_from_categories (NVCategory.cu:430-514): one category over the concatenation of the inputs'
    rows, built from their keys and values only (no pass over strings)."""
    arr = (C.c_void_p * len(cats))(*[c.m_cptr for c in cats])
    return nvcategory(check_handle(lib().custr_category_merge(arr, len(cats), 1), "from_categories"))


class nvcategory:
    """Dictionary-encoded strings: sorted unique keys + int32 values.  reference nvcategory.py:136"""

    def __init__(self, cptr):
        self.m_cptr = cptr

    def __del__(self):
        if getattr(self, "m_cptr", None):
            try:
                lib().custr_category_free(self.m_cptr)
            except Exception:
                pass
            self.m_cptr = 0

    def __repr__(self):
        return "<nvcategory keys={},values={}>".format(self.keys_size(), self.size())

    def size(self):
        return int(lib().custr_category_size(self.m_cptr))

    def keys_size(self):
        return int(lib().custr_category_keys_size(self.m_cptr))

    def keys(self):
        """nvstrings of the sorted distinct keys.  reference nvcategory.py:242 -> NVCategory.cu:724-750"""
        return _nvs.nvstrings(check_handle(lib().custr_category_keys(self.m_cptr), "keys"))

    def values(self, devptr=0):
        """int32 key index per row.  reference nvcategory.py:364 -> NVCategory.cu:866-878"""
        if devptr:
            check_rc(lib().custr_category_values(self.m_cptr, as_ptr(devptr), 1), "values")
            return devptr
        out = np.zeros(max(self.size(), 1), np.int32)
        check_rc(lib().custr_category_values(self.m_cptr, as_ptr(out), 0), "values")
        return out[: self.size()].tolist()

    def values_cpointer(self):
        """Device pointer to the int32 values.  reference nvcategory.py:343"""
        return int(lib().custr_category_values_cptr(self.m_cptr) or 0)

    def to_strings(self):
        """reference nvcategory.py:492"""
        return self.keys().gather(self.values_cpointer(), self.size())

    def merge_and_remap(self, other):
        """New category over this + other: keys = sorted union, values = both value arrays remapped and appended.
        reference nvcategory.py:merge_and_remap -> NVCategory.cu:1339 (create_from_categories :430)"""
        return from_categories([self, other])

    def merge_category(self, other):
        """New category: this one's keys followed by other's new keys (not re-sorted), this one's values unchanged,
        other's remapped and appended.  reference nvcategory.py:merge_category -> NVCategory.cu:1223"""
        arr = (C.c_void_p * 2)(self.m_cptr, other.m_cptr)
        return nvcategory(check_handle(lib().custr_category_merge(arr, 2, 0), "merge_category"))

    def _keys_op(self, strs, op, what):
        if strs is not None and type(strs).__name__ != "nvstrings":
            strs = _nvs.to_device(list(strs))
        h = lib().custr_category_keys_op(self.m_cptr, strs.m_cptr if strs is not None else None, op)
        return nvcategory(check_handle(h, what))

    def add_keys(self, strs, nulls=None):
        """keys = sorted union with strs, values remapped.  reference nvcategory.py:720 -> NVCategory.cu:1375"""
        return self._keys_op(strs, 0, "add_keys")

    def remove_keys(self, strs, nulls=None):
        """keys minus strs; values of removed keys become -1.  reference nvcategory.py:750 -> NVCategory.cu:1482"""
        return self._keys_op(strs, 1, "remove_keys")

    def set_keys(self, strs, nulls=None):
        """keys = sorted distinct strs; values of keys that are gone become -1.  reference nvcategory.py:805 -> NVCategory.cu:1708"""
        return self._keys_op(strs, 2, "set_keys")

    def remove_unused_keys(self):
        """reference nvcategory.py:780 -> NVCategory.cu:1567"""
        return self._keys_op(None, 3, "remove_unused_keys")

    def _positions(self, indexes, count):
        if isinstance(indexes, (list, tuple)):
            indexes = np.asarray(indexes, np.int32)
        if isinstance(indexes, np.ndarray):
            a = np.ascontiguousarray(indexes, np.int32)
            return a, as_ptr(a), len(a), 0
        return indexes, as_ptr(indexes), count, 1

    def gather(self, indexes, count=0):
        """same keys, values = indexes.  reference nvcategory.py:630 -> NVCategory.cu:1142"""
        keep, p, n, dev = self._positions(indexes, count)
        return nvcategory(check_handle(lib().custr_category_gather(self.m_cptr, p, n, dev, 0), "gather"))

    def gather_and_remap(self, indexes, count=0):
        """keys = the keys the indexes use, values remapped.  reference nvcategory.py:590 -> NVCategory.cu:1084"""
        keep, p, n, dev = self._positions(indexes, count)
        return nvcategory(check_handle(lib().custr_category_gather(self.m_cptr, p, n, dev, 1), "gather_and_remap"))

    def gather_strings(self, indexes, count=0):
        """the key strings at the given key indexes.  reference nvcategory.py:518 -> NVCategory.cu:1011"""
        keep, p, n, dev = self._positions(indexes, count)
        return _nvs.nvstrings(check_handle(lib().custr_category_gather_strings(self.m_cptr, p, n, dev), "gather_strings"))

    def __getattr__(self, name):
        if name.startswith("_") or name == "m_cptr":
            raise AttributeError(name)
        raise NotImplementedError("nvcategory.%s is outside the hot path implemented by custrings_b200 (SURVEY.md section 8)" % name)

"""ctypes binding of libcustr.so (include/custr.h).  This replaces the reference's CPython extension modules
pyniNVStrings / pyniNVCategory / pyniNVText (python/cpp/pystrings.cpp, pycategory.cpp, pytext.cpp): same role,
but it binds the thin C-ABI instead of the C++ classes.

There is NO CPU fallback: if the shared library is missing or cannot be loaded this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CUSTR_LIB") or os.path.join(_HERE, "libcustr.so")  # CUSTR_LIB: A/B variant builds (tools/build_variant.py)
_lib = None

vp, ci, cu, cp, cl = C.c_void_p, C.c_int, C.c_uint, C.c_char_p, C.c_longlong

_SIGNATURES = {
    # name: (restype, argtypes)
    "custr_last_error": (cp, []),
    "custr_version": (cp, []),
    "custr_set_device": (ci, [ci]),
    "custr_set_stream": (None, [vp]),
    "custr_sync": (ci, []),
    "custr_launch_count": (cl, []),
    "custr_last_regex_tier": (cp, []),
    "custr_set_regex_tier": (None, [ci]),
    "custr_set_item_kib": (None, [ci]),
    "custr_set_jit": (None, [ci, cl]),
    "custr_jit_launch_count": (cl, []),
    "custr_jit_note": (cp, []),
    "custr_set_profiling": (None, [ci]),
    "custr_last_kernel_ms": (C.c_float, []),
    "custr_create_from_offsets": (vp, [vp, ci, vp, vp, ci, ci]),
    "custr_adopt_device": (vp, [vp, ci, vp, vp, ci]),
    "custr_create_from_array": (vp, [vp, cu]),
    "custr_create_from_index": (vp, [vp, cu, ci, ci]),
    "custr_ipc_export": (ci, [vp, vp]),
    "custr_ipc_import": (vp, [vp]),
    "custr_column_free": (None, [vp]),
    "custr_release_cached_memory": (None, []),
    "custr_size": (cu, [vp]),
    "custr_chars_bytes": (cl, [vp]),
    "custr_null_count": (ci, [vp]),
    "custr_chars_ptr": (vp, [vp]),
    "custr_offsets_ptr": (vp, [vp]),
    "custr_validity_ptr": (vp, [vp]),
    "custr_create_offsets": (ci, [vp, vp, vp, vp, ci]),
    "custr_set_null_bitarray": (ci, [vp, vp, ci, ci]),
    "custr_byte_count": (cl, [vp, vp, ci]),
    "custr_len": (ci, [vp, vp, ci]),
    "custr_hash": (ci, [vp, vp, ci]),
    "custr_contains_re": (ci, [vp, cp, vp, ci]),
    "custr_match": (ci, [vp, cp, vp, ci]),
    "custr_count_re": (ci, [vp, cp, vp, ci]),
    "custr_replace_re": (vp, [vp, cp, cp, ci]),
    "custr_replace_re_multi": (vp, [vp, vp, ci, vp]),
    "custr_replace_with_backrefs": (vp, [vp, cp, cp]),
    "custr_regex_describe": (ci, [cp, vp, C.c_size_t]),
    "custr_findall": (ci, [vp, cp, vp, ci]),
    "custr_findall_record": (ci, [vp, cp, vp, vp, ci]),
    "custr_extract": (ci, [vp, cp, vp, ci]),
    "custr_find": (ci, [vp, cp, ci, ci, vp, ci]),
    "custr_rfind": (ci, [vp, cp, ci, ci, vp, ci]),
    "custr_contains": (ci, [vp, cp, vp, ci]),
    "custr_startswith": (ci, [vp, cp, vp, ci]),
    "custr_endswith": (ci, [vp, cp, vp, ci]),
    "custr_find_multiple": (ci, [vp, vp, vp, ci]),
    "custr_replace": (vp, [vp, cp, cp, ci]),
    "custr_replace_multi": (vp, [vp, vp, vp]),
    "custr_split": (ci, [vp, cp, ci, vp, ci]),
    "custr_rsplit": (ci, [vp, cp, ci, vp, ci]),
    "custr_split_record": (ci, [vp, cp, ci, vp, vp, ci]),
    "custr_rsplit_record": (ci, [vp, cp, ci, vp, vp, ci]),
    "custr_partition": (vp, [vp, cp, ci]),
    "custr_find_from": (ci, [vp, cp, vp, vp, vp, ci]),
    "custr_match_strings": (ci, [vp, vp, vp, ci]),
    "custr_slice_rows": (vp, [vp, ci, ci]),
    "custr_gather": (vp, [vp, vp, ci, ci]),
    "custr_tokenize": (vp, [vp, cp]),
    "custr_token_count": (ci, [vp, cp, vp, ci]),
    "custr_category_create": (vp, [vp, ci]),
    "custr_category_free": (None, [vp]),
    "custr_category_size": (cu, [vp]),
    "custr_category_keys_size": (cu, [vp]),
    "custr_category_keys": (vp, [vp]),
    "custr_category_values": (ci, [vp, vp, ci]),
    "custr_category_values_cptr": (vp, [vp]),
    "custr_category_remap_to_union": (vp, [vp, vp]),
    "custr_category_merge": (vp, [vp, ci, ci]),
    "custr_is_class": (ci, [vp, ci, vp, ci]),
    "custr_case": (vp, [vp, ci]),
    "custr_strip": (vp, [vp, cp, ci]),
    "custr_slice": (vp, [vp, ci, ci, ci]),
    "custr_category_keys_op": (vp, [vp, vp, ci]),
    "custr_category_gather": (vp, [vp, vp, ci, ci, ci]),
    "custr_category_gather_strings": (vp, [vp, vp, ci, ci]),
    "custr_comm_unique_id": (ci, [vp]),
    "custr_comm_create": (vp, [ci, ci, vp]),
    "custr_comm_destroy": (None, [vp]),
    "custr_category_create_sharded": (vp, [vp, vp, vp]),
}

EXPORTED_SYMBOLS = sorted(_SIGNATURES)

CUSTR_ERR_ARG, CUSTR_ERR_INVALID, CUSTR_ERR_CUDA, CUSTR_ERR_ALLOC = -1, -2, -3, -4


def lib():
    """Load libcustr.so; raises (never falls back) when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "custrings_b200: %s is missing - build it with `python -m custrings_b200.build` "
                "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError here == header/library mismatch: fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error():
    return lib().custr_last_error().decode("utf-8", "replace")


def check_handle(h, what):
    """NULL handle -> ValueError like the reference binding (pystrings.cpp:1915-1932)."""
    if not h:
        raise ValueError("%s: %s" % (what, last_error()))
    return h


def check_rc(rc, what):
    """Negative status other than the reference's plain -1 -> ValueError / RuntimeError."""
    if rc is not None and rc <= CUSTR_ERR_INVALID:
        msg = "%s: %s" % (what, last_error())
        if rc in (CUSTR_ERR_CUDA, CUSTR_ERR_ALLOC):
            raise RuntimeError(msg)
        raise ValueError(msg)
    return rc


def as_ptr(x):
    """int | numpy array | torch tensor | numba DeviceNDArray | ctypes -> raw address (reference DataBuffer<T>,
    pystrings.cpp:43-175)."""
    if x is None:
        return None
    if isinstance(x, int):
        return x or None
    if hasattr(x, "data_ptr"):  # torch
        return x.data_ptr() or None
    if hasattr(x, "__cuda_array_interface__"):
        return x.__cuda_array_interface__["data"][0] or None
    if hasattr(x, "device_ctypes_pointer"):
        return x.device_ctypes_pointer.value
    if hasattr(x, "ctypes"):  # numpy
        return x.ctypes.data or None
    if isinstance(x, (C.c_void_p,)):
        return x.value
    return C.cast(x, C.c_void_p).value

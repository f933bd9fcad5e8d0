"""ctypes view of baseline/_ref/libref_gpu.so — the UNMODIFIED reference custrings CUDA sources built for sm_100
(baseline/Makefile).  Baseline infrastructure: used by bench.py's `reference_gpu` leg and tools/bench_ref_gpu.py only."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libref_gpu.so")
_lib = None


def available():
    return os.path.exists(_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_PATH)
        vp, ip, cp = C.c_void_p, C.c_int, C.c_char_p
        L.refgpu_last_error.restype = cp
        L.refgpu_create.restype = vp
        L.refgpu_create.argtypes = [vp, ip, vp, vp, ip]
        L.refgpu_destroy.argtypes = [vp]
        L.refgpu_size.argtypes = [vp]
        L.refgpu_size.restype = C.c_uint
        for name in ("refgpu_contains_re", "refgpu_count_re", "refgpu_contains"):
            f = getattr(L, name)
            f.restype = ip
            f.argtypes = [vp, cp, vp]
        L.refgpu_replace_re.restype = vp
        L.refgpu_replace_re.argtypes = [vp, cp, cp, ip]
        L.refgpu_replace.restype = vp
        L.refgpu_replace.argtypes = [vp, cp, cp, ip]
        L.refgpu_split.restype = ip
        L.refgpu_split.argtypes = [vp, cp, ip, C.POINTER(vp), ip]
        L.refgpu_split_record_total.restype = C.c_long
        L.refgpu_split_record_total.argtypes = [vp, cp, ip]
        L.refgpu_tokenize.restype = vp
        L.refgpu_tokenize.argtypes = [vp, cp]
        L.refgpu_category.restype = vp
        L.refgpu_category.argtypes = [vp]
        L.refgpu_category_destroy.argtypes = [vp]
        L.refgpu_category_keys_size.restype = C.c_uint
        L.refgpu_category_keys_size.argtypes = [vp]
        L.refgpu_category_values.restype = ip
        L.refgpu_category_values.argtypes = [vp, vp]
        L.refgpu_export.restype = ip
        L.refgpu_export.argtypes = [vp, vp, vp, vp]
        L.refgpu_memsize.restype = C.c_long
        L.refgpu_memsize.argtypes = [vp]
        _lib = L
    return _lib


class RefGpuStrings:
    """NVStrings instance of the reference CUDA build; inputs are torch CUDA tensors."""

    def __init__(self, h):
        if not h:
            raise RuntimeError("reference GPU build: " + lib().refgpu_last_error().decode())
        self.h = h

    @classmethod
    def from_device(cls, d_chars, d_offsets, n, d_validity=None, nulls=0):
        return cls(lib().refgpu_create(d_chars.data_ptr(), n, d_offsets.data_ptr(), d_validity.data_ptr() if d_validity is not None and nulls else None, nulls))

    def size(self):
        return lib().refgpu_size(self.h)

    def contains_re(self, pat, d_out):
        return lib().refgpu_contains_re(self.h, pat.encode(), d_out.data_ptr())

    def count_re(self, pat, d_out):
        return lib().refgpu_count_re(self.h, pat.encode(), d_out.data_ptr())

    def replace_re(self, pat, repl, maxrepl=-1):
        return RefGpuStrings(lib().refgpu_replace_re(self.h, pat.encode(), repl.encode(), maxrepl))

    def replace(self, tgt, repl, maxrepl=-1):
        return RefGpuStrings(lib().refgpu_replace(self.h, tgt.encode(), repl.encode(), maxrepl))

    def split(self, delim, maxsplit=-1, cap=64):
        arr = (C.c_void_p * cap)()
        k = lib().refgpu_split(self.h, delim.encode(), maxsplit, arr, cap)
        if k < 0:
            raise RuntimeError(lib().refgpu_last_error().decode())
        return [RefGpuStrings(arr[i]) for i in range(min(k, cap))]

    def split_record_total(self, delim, maxsplit=-1):
        return lib().refgpu_split_record_total(self.h, delim.encode(), maxsplit)

    def tokenize(self, delim=None):
        return RefGpuStrings(lib().refgpu_tokenize(self.h, delim.encode() if delim else None))

    def category(self):
        c = lib().refgpu_category(self.h)
        if not c:
            raise RuntimeError(lib().refgpu_last_error().decode())
        return c

    def memsize(self):
        return lib().refgpu_memsize(self.h)

    def free(self):
        if self.h:
            lib().refgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

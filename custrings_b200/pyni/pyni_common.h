// Shared helpers of the CPython extension modules pyniNVStrings / pyniNVCategory / pyniNVText — this repo's replacement for
// the reference's binding layer (python/cpp/pystrings.cpp:3860-3973, pycategory.cpp:900-937, pytext.cpp:653-674) for the
// hot-path subset: same module names, same n_* function names and positional arguments, handles as Python ints
// (PyLong_AsVoidPtr), GIL released around every library call, failures -> ValueError + None (pystrings.cpp:1915-1932,
// 2599-2618), host results as lists with None for null rows (:2644-2664).  Underneath sits libcustr.so's C-ABI
// (include/custr.h); a handle is a custr_column* / custr_category*.
#pragma once
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <cstdint>
#include <string>
#include <vector>
#include "../../include/custr.h"

namespace pyni {

inline void* ptr_arg(PyObject* args, int i)
{
    PyObject* o = PyTuple_GetItem(args, i);
    if (!o || o == Py_None) return nullptr;
    if (PyLong_Check(o)) return PyLong_AsVoidPtr(o);
    // objects that expose a device pointer: __cuda_array_interface__ (numba / cupy / torch), or data_ptr() (torch)
    if (PyObject_HasAttrString(o, "__cuda_array_interface__")) {
        PyObject* d = PyObject_GetAttrString(o, "__cuda_array_interface__");
        PyObject* data = d ? PyDict_GetItemString(d, "data") : nullptr;
        void* p = data ? PyLong_AsVoidPtr(PyTuple_GetItem(data, 0)) : nullptr;
        Py_XDECREF(d);
        return p;
    }
    if (PyObject_HasAttrString(o, "data_ptr")) {
        PyObject* r = PyObject_CallMethod(o, "data_ptr", nullptr);
        void* p = r ? PyLong_AsVoidPtr(r) : nullptr;
        Py_XDECREF(r);
        return p;
    }
    return nullptr;
}
inline const custr_column* col_arg(PyObject* args, int i) { return (const custr_column*)ptr_arg(args, i); }
// str -> UTF-8 (borrowed), None -> nullptr
inline const char* str_arg(PyObject* args, int i)
{
    PyObject* o = PyTuple_GetItem(args, i);
    if (!o || o == Py_None) return nullptr;
    return PyUnicode_AsUTF8(o);
}
inline long long int_arg(PyObject* args, int i, long long dflt)
{
    PyObject* o = i < PyTuple_Size(args) ? PyTuple_GetItem(args, i) : nullptr;
    if (!o || o == Py_None) return dflt;
    return PyLong_AsLongLong(o);
}
inline bool true_arg(PyObject* args, int i)
{
    PyObject* o = i < PyTuple_Size(args) ? PyTuple_GetItem(args, i) : nullptr;
    return o && PyObject_IsTrue(o) == 1;
}
// the reference raises ValueError(message) and returns None (pystrings.cpp:2599-2618)
inline PyObject* fail_none()
{
    const char* m = custr_last_error();
    PyErr_SetString(PyExc_ValueError, (m && *m) ? m : "custrings: call failed");
    return nullptr;
}
inline PyObject* handle_or_none(const void* h)
{
    if (!h) return fail_none();
    return PyLong_FromVoidPtr((void*)h);
}
// validity of every row as host bytes (1 = valid)
inline std::vector<uint8_t> valid_rows(const custr_column* c)
{
    const uint32_t n = custr_size(c);
    std::vector<uint8_t> bits((n + 7) / 8 + 1, 0xff), out(n, 1);
    int nulls = 0;
    Py_BEGIN_ALLOW_THREADS
    nulls = custr_set_null_bitarray(c, bits.data(), 0, 0);
    Py_END_ALLOW_THREADS
    if (nulls > 0)
        for (uint32_t i = 0; i < n; ++i) out[i] = (bits[i >> 3] >> (i & 7)) & 1;
    return out;
}
// Per-row results: into the caller's device array when devptr != 0 (returns it), else a list with None for null rows.
// call(dst, devmem) -> status (negative = error).  kind: 'b' bool from uint8, 'i' int32 (null rows are those < null_below),
// 'u' uint32 (never null)
template <typename T, typename F>
PyObject* row_results(const custr_column* c, void* devptr, char kind, int null_below, F call)
{
    if (!c) { PyErr_SetString(PyExc_ValueError, "custrings: null handle"); return nullptr; }
    long long rc = 0;
    if (devptr) {
        Py_BEGIN_ALLOW_THREADS
        rc = call((T*)devptr, 1);
        Py_END_ALLOW_THREADS
        if (rc <= CUSTR_ERR_INVALID) return fail_none();
        if (rc < 0) Py_RETURN_NONE;
        return PyLong_FromVoidPtr(devptr);
    }
    const uint32_t n = custr_size(c);
    if (n == 0) return PyList_New(0);
    std::vector<T> host(n);
    Py_BEGIN_ALLOW_THREADS
    rc = call(host.data(), 0);
    Py_END_ALLOW_THREADS
    if (rc <= CUSTR_ERR_INVALID) return fail_none();
    if (rc < 0) Py_RETURN_NONE;
    std::vector<uint8_t> valid;
    if (kind == 'b') valid = valid_rows(c);
    PyObject* list = PyList_New(n);
    for (uint32_t i = 0; i < n; ++i) {
        PyObject* v;
        if (kind == 'b') v = valid[i] ? PyBool_FromLong(host[i] != 0) : (Py_INCREF(Py_None), Py_None);
        else if (kind == 'i') v = ((long long)host[i] < null_below) ? (Py_INCREF(Py_None), Py_None) : PyLong_FromLongLong((long long)host[i]);
        else v = PyLong_FromUnsignedLongLong((unsigned long long)host[i]);
        PyList_SetItem(list, i, v);
    }
    return list;
}
// list of new column handles (0 stays 0: the shims skip / map it to None)
inline PyObject* handle_list(const std::vector<custr_column*>& cols)
{
    PyObject* list = PyList_New((Py_ssize_t)cols.size());
    for (size_t i = 0; i < cols.size(); ++i) PyList_SetItem(list, (Py_ssize_t)i, PyLong_FromVoidPtr((void*)cols[i]));
    return list;
}
// the strings of a column as a Python list (None for null rows): n_createHostStrings
inline PyObject* host_strings(const custr_column* c)
{
    const uint32_t n = custr_size(c);
    if (n == 0) return PyList_New(0);
    const long long bytes = custr_chars_bytes(c);
    std::vector<char> chars((size_t)(bytes > 0 ? bytes : 1));
    std::vector<int32_t> off(n + 1);
    std::vector<uint8_t> bits((n + 7) / 8 + 1, 0);
    int rc = 0;
    Py_BEGIN_ALLOW_THREADS
    rc = custr_create_offsets(c, chars.data(), off.data(), bits.data(), 0);
    Py_END_ALLOW_THREADS
    if (rc < 0) return fail_none();
    const bool has_nulls = custr_null_count(c) > 0;
    PyObject* list = PyList_New(n);
    for (uint32_t i = 0; i < n; ++i) {
        if (has_nulls && !((bits[i >> 3] >> (i & 7)) & 1)) { Py_INCREF(Py_None); PyList_SetItem(list, i, Py_None); continue; }
        PyList_SetItem(list, i, PyUnicode_DecodeUTF8(chars.data() + off[i], off[i + 1] - off[i], "replace"));
    }
    return list;
}
// Python list of str / None -> new column: n_createFromHostStrings
inline custr_column* column_from_list(PyObject* list)
{
    if (!PyList_Check(list)) {
        if (PyUnicode_Check(list)) {
            const char* one[1] = {PyUnicode_AsUTF8(list)};
            return custr_create_from_array(one, 1);
        }
        PyErr_SetString(PyExc_ValueError, "nvstrings: expected a list of strings");
        return nullptr;
    }
    const Py_ssize_t n = PyList_Size(list);
    std::vector<const char*> ptrs((size_t)n, nullptr);
    for (Py_ssize_t i = 0; i < n; ++i) {
        PyObject* o = PyList_GetItem(list, i);
        if (o != Py_None && PyUnicode_Check(o)) ptrs[(size_t)i] = PyUnicode_AsUTF8(o);
    }
    custr_column* c = nullptr;
    Py_BEGIN_ALLOW_THREADS
    c = custr_create_from_array(ptrs.data(), (uint32_t)n);
    Py_END_ALLOW_THREADS
    return c;
}
}  // namespace pyni

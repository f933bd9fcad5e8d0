// pyniNVStrings — CPython extension module, hot-path subset of the reference's python/cpp/pystrings.cpp (method table
// :3860-3973) over libcustr.so's C-ABI.  Function names, positional arguments and result conventions follow the reference
// function cited at each entry, so the reference's python/nvstrings.py binds these unmodified for the methods listed here.
#include "pyni_common.h"
using namespace pyni;

#define GIL_FREE(stmt) Py_BEGIN_ALLOW_THREADS stmt; Py_END_ALLOW_THREADS

// pystrings.cpp:291-375
static PyObject* n_createFromHostStrings(PyObject*, PyObject* args) { return handle_or_none(column_from_list(PyTuple_GetItem(args, 0))); }
// :2-? n_destroyStrings
static PyObject* n_destroyStrings(PyObject*, PyObject* args)
{
    custr_column* c = (custr_column*)ptr_arg(args, 0);
    GIL_FREE(custr_column_free(c));
    return PyLong_FromLong(0);
}
// :450-520 to_host
static PyObject* n_createHostStrings(PyObject*, PyObject* args) { return host_strings(col_arg(args, 0)); }

struct BufArg {  // int address | buffer-protocol object | None (the reference's handling in n_createFromOffsets, :377-447)
    Py_buffer view{};
    bool held = false;
    void* ptr = nullptr;
    BufArg(PyObject* o)
    {
        if (!o || o == Py_None) return;
        if (PyLong_Check(o)) ptr = PyLong_AsVoidPtr(o);
        else if (PyObject_CheckBuffer(o) && PyObject_GetBuffer(o, &view, PyBUF_SIMPLE) == 0) { held = true; ptr = view.buf; }
        else {
            PyErr_Clear();
            PyObject* t = PyTuple_Pack(1, o);
            ptr = ptr_arg(t, 0);
            Py_DECREF(t);
        }
    }
    ~BufArg() { if (held) PyBuffer_Release(&view); }
};
// :377-447  (sbuf, obuf, scount, nbuf, ncount, bdevmem)
static PyObject* n_createFromOffsets(PyObject*, PyObject* args)
{
    if (PyTuple_GetItem(args, 0) == Py_None || PyTuple_GetItem(args, 1) == Py_None) {
        PyErr_SetString(PyExc_ValueError, "nvstrings: missing parameter");
        return nullptr;
    }
    BufArg s(PyTuple_GetItem(args, 0)), o(PyTuple_GetItem(args, 1)), nb(PyTuple_GetItem(args, 3));
    const int scount = (int)int_arg(args, 2, 0), ncount = nb.ptr ? (int)int_arg(args, 4, 0) : 0;
    const int devmem = true_arg(args, 5);
    custr_column* c = nullptr;
    GIL_FREE(c = custr_create_from_offsets((const char*)s.ptr, scount, (const int32_t*)o.ptr, (const uint8_t*)nb.ptr, ncount, devmem));
    return handle_or_none(c);
}
// n_create_offsets (cptr, sbuf, obuf, nbuf, bdevmem) -> null count
static PyObject* n_create_offsets(PyObject*, PyObject* args)
{
    BufArg s(PyTuple_GetItem(args, 1)), o(PyTuple_GetItem(args, 2)), nb(PyTuple_GetItem(args, 3));
    int rc = 0;
    const custr_column* c = col_arg(args, 0);
    const int devmem = true_arg(args, 4);
    GIL_FREE(rc = custr_create_offsets(c, (char*)s.ptr, (int32_t*)o.ptr, (uint8_t*)nb.ptr, devmem));
    if (rc <= CUSTR_ERR_INVALID) return fail_none();
    return PyLong_FromLong(rc);
}
static PyObject* n_size(PyObject*, PyObject* args) { return PyLong_FromUnsignedLong(custr_size(col_arg(args, 0))); }
// :911-945
static PyObject* n_len(PyObject*, PyObject* args)
{
    const custr_column* c = col_arg(args, 0);
    return row_results<int32_t>(c, ptr_arg(args, 1), 'i', 0, [&](int32_t* d, int dm) { return (long long)custr_len(c, d, dm); });
}
// :947-957 (cptr, vals, bdevmem) -> total bytes
static PyObject* n_byte_count(PyObject*, PyObject* args)
{
    long long rc = 0;
    const custr_column* c = col_arg(args, 0);
    BufArg vals(PyTuple_GetItem(args, 1));
    const int devmem = true_arg(args, 2);
    GIL_FREE(rc = custr_byte_count(c, (int32_t*)vals.ptr, devmem));
    return PyLong_FromLongLong(rc);
}
static PyObject* n_set_null_bitmask(PyObject*, PyObject* args)
{
    int rc = 0;
    const custr_column* c = col_arg(args, 0);
    BufArg nb(PyTuple_GetItem(args, 1));
    const int devmem = true_arg(args, 2);
    GIL_FREE(rc = custr_set_null_bitarray(c, (uint8_t*)nb.ptr, 0, devmem));
    return PyLong_FromLong(rc);
}
static PyObject* n_null_count(PyObject*, PyObject* args)
{
    const custr_column* c = col_arg(args, 0);
    int rc = 0;
    if (true_arg(args, 1)) { GIL_FREE(rc = custr_set_null_bitarray(c, nullptr, 1, 0)); }
    else rc = custr_null_count(c);
    return PyLong_FromLong(rc);
}
static PyObject* n_hash(PyObject*, PyObject* args)
{
    const custr_column* c = col_arg(args, 0);
    return row_results<uint32_t>(c, ptr_arg(args, 1), 'u', 0, [&](uint32_t* d, int dm) { return (long long)custr_hash(c, d, dm); });
}
static PyObject* n_copy(PyObject*, PyObject* args)
{
    const custr_column* c = col_arg(args, 0);
    custr_column* r = nullptr;
    GIL_FREE(r = custr_slice_rows(c, 0, (int32_t)custr_size(c)));
    return handle_or_none(r);
}
// n_gather (cptr, indexes, count): list | buffer | device pointer (DataBuffer<int>, :43-175)
static PyObject* n_gather(PyObject*, PyObject* args)
{
    const custr_column* c = col_arg(args, 0);
    PyObject* idx = PyTuple_GetItem(args, 1);
    custr_column* r = nullptr;
    if (PyList_Check(idx)) {
        std::vector<int32_t> v((size_t)PyList_Size(idx));
        for (size_t i = 0; i < v.size(); ++i) v[i] = (int32_t)PyLong_AsLong(PyList_GetItem(idx, (Py_ssize_t)i));
        GIL_FREE(r = custr_gather(c, v.data(), (int32_t)v.size(), 0));
    } else if (!PyLong_Check(idx) && PyObject_CheckBuffer(idx)) {
        BufArg b(idx);
        const int32_t count = (int32_t)(b.view.len / 4);
        GIL_FREE(r = custr_gather(c, (const int32_t*)b.ptr, count, 0));
    } else {
        void* d = ptr_arg(args, 1);
        const int32_t count = (int32_t)int_arg(args, 2, 0);
        GIL_FREE(r = custr_gather(c, (const int32_t*)d, count, 1));
    }
    return handle_or_none(r);
}

// ---- regex: n_contains :2588-2664 (cptr, pat, regex, devptr), n_match (cptr, pat, devptr), n_count (cptr, pat, devptr)
static PyObject* n_contains(PyObject*, PyObject* args)
{
    const custr_column* c = col_arg(args, 0);
    const char* pat = str_arg(args, 1);
    const bool regex = true_arg(args, 2);
    return row_results<uint8_t>(c, ptr_arg(args, 3), 'b', 0, [&](uint8_t* d, int dm) {
        return (long long)(regex ? custr_contains_re(c, pat, d, dm) : custr_contains(c, pat, d, dm));
    });
}
static PyObject* n_match(PyObject*, PyObject* args)
{
    const custr_column* c = col_arg(args, 0);
    const char* pat = str_arg(args, 1);
    return row_results<uint8_t>(c, ptr_arg(args, 2), 'b', 0, [&](uint8_t* d, int dm) { return (long long)custr_match(c, pat, d, dm); });
}
static PyObject* n_count(PyObject*, PyObject* args)
{
    const custr_column* c = col_arg(args, 0);
    const char* pat = str_arg(args, 1);
    void* devptr = ptr_arg(args, 2);
    if (devptr) return row_results<int32_t>(c, devptr, 'i', 0, [&](int32_t* d, int dm) { return (long long)custr_count_re(c, pat, d, dm); });
    // host list: null rows are None (count_re reports 0 for them: take the validity instead)
    const uint32_t n = custr_size(c);
    if (n == 0) return PyList_New(0);
    std::vector<int32_t> host(n);
    int rc = 0;
    GIL_FREE(rc = custr_count_re(c, pat, host.data(), 0));
    if (rc <= CUSTR_ERR_INVALID) return fail_none();
    if (rc < 0) Py_RETURN_NONE;
    const std::vector<uint8_t> valid = valid_rows(c);
    PyObject* list = PyList_New(n);
    for (uint32_t i = 0; i < n; ++i) PyList_SetItem(list, i, valid[i] ? PyLong_FromLong(host[i]) : (Py_INCREF(Py_None), Py_None));
    return list;
}
// n_replace (cptr, pat, repl, n, regex)
static PyObject* n_replace(PyObject*, PyObject* args)
{
    const custr_column* c = col_arg(args, 0);
    const char *pat = str_arg(args, 1), *repl = str_arg(args, 2);
    const int n = (int)int_arg(args, 3, -1);
    const bool regex = true_arg(args, 4);
    custr_column* r = nullptr;
    GIL_FREE(r = regex ? custr_replace_re(c, pat, repl, n) : custr_replace(c, pat, repl, n));
    return handle_or_none(r);
}
// n_replace_multi (cptr, pats list, repls cptr, regex)
static PyObject* n_replace_multi(PyObject*, PyObject* args)
{
    const custr_column* c = col_arg(args, 0);
    PyObject* pats = PyTuple_GetItem(args, 1);
    const custr_column* repls = col_arg(args, 2);
    const bool regex = true_arg(args, 3);
    custr_column* r = nullptr;
    if (regex) {
        if (!PyList_Check(pats)) { PyErr_SetString(PyExc_ValueError, "replace_multi: patterns must be a list of strings"); return nullptr; }
        std::vector<const char*> p((size_t)PyList_Size(pats));
        for (size_t i = 0; i < p.size(); ++i) p[i] = PyUnicode_AsUTF8(PyList_GetItem(pats, (Py_ssize_t)i));
        GIL_FREE(r = custr_replace_re_multi(c, p.data(), (int32_t)p.size(), repls));
        return handle_or_none(r);
    }
    // literal targets: a list of strings or an nvstrings handle
    custr_column* targets = nullptr;
    bool own = false;
    if (PyList_Check(pats)) { targets = column_from_list(pats); own = true; }
    else targets = (custr_column*)ptr_arg(args, 1);
    if (!targets) return fail_none();
    GIL_FREE(r = custr_replace_multi(c, targets, repls));
    if (own) custr_column_free(targets);
    return handle_or_none(r);
}
static PyObject* n_replace_with_backrefs(PyObject*, PyObject* args)
{
    const custr_column* c = col_arg(args, 0);
    const char *pat = str_arg(args, 1), *repl = str_arg(args, 2);
    custr_column* r = nullptr;
    GIL_FREE(r = custr_replace_with_backrefs(c, pat, repl));
    return handle_or_none(r);
}

// ---- find family: n_find :2192-2235 (cptr, sub, start, end|None, devptr); -1 not found, null rows None (values < -1)
template <typename F>
static PyObject* find_like(PyObject* args, F fn)
{
    const custr_column* c = col_arg(args, 0);
    const char* sub = str_arg(args, 1);
    const int start = (int)int_arg(args, 2, 0), end = (int)int_arg(args, 3, -1);
    return row_results<int32_t>(c, ptr_arg(args, 4), 'i', -1, [&](int32_t* d, int dm) { return (long long)fn(c, sub, start, end, d, dm); });
}
static PyObject* n_find(PyObject*, PyObject* args) { return find_like(args, custr_find); }
static PyObject* n_rfind(PyObject*, PyObject* args) { return find_like(args, custr_rfind); }
// n_find_from (cptr, sub, starts devptr, ends devptr, devptr)
static PyObject* n_find_from(PyObject*, PyObject* args)
{
    const custr_column* c = col_arg(args, 0);
    const char* sub = str_arg(args, 1);
    const int32_t *starts = (const int32_t*)ptr_arg(args, 2), *ends = (const int32_t*)ptr_arg(args, 3);
    void* devptr = ptr_arg(args, 4);
    if (devptr) return row_results<int32_t>(c, devptr, 'i', -1, [&](int32_t* d, int dm) { return (long long)custr_find_from(c, sub, starts, ends, d, dm); });
    // host results with device start / end arrays: compute on the device, copy back through a gather-free path
    PyErr_SetString(PyExc_ValueError, "find_from: pass devptr (the start / end arrays are device arrays)");
    return nullptr;
}
template <typename F>
static PyObject* bool_by_str(PyObject* args, F fn)
{
    const custr_column* c = col_arg(args, 0);
    const char* s = str_arg(args, 1);
    return row_results<uint8_t>(c, ptr_arg(args, 2), 'b', 0, [&](uint8_t* d, int dm) { return (long long)fn(c, s, d, dm); });
}
static PyObject* n_startswith(PyObject*, PyObject* args) { return bool_by_str(args, custr_startswith); }
static PyObject* n_endswith(PyObject*, PyObject* args) { return bool_by_str(args, custr_endswith); }
// n_match_strings (cptr, strs: list | cptr, devptr)
static PyObject* n_match_strings(PyObject*, PyObject* args)
{
    const custr_column* c = col_arg(args, 0);
    PyObject* o = PyTuple_GetItem(args, 1);
    custr_column* other = nullptr;
    bool own = false;
    if (PyList_Check(o)) { other = column_from_list(o); own = true; }
    else other = (custr_column*)ptr_arg(args, 1);
    if (!other) return fail_none();
    PyObject* r = row_results<uint8_t>(c, ptr_arg(args, 2), 'b', 0, [&](uint8_t* d, int dm) { return (long long)custr_match_strings(c, other, d, dm); });
    if (own) custr_column_free(other);
    return r;
}
// n_find_multiple (cptr, strs: list | cptr, devptr) -> devptr | list of per-row lists
static PyObject* n_find_multiple(PyObject*, PyObject* args)
{
    const custr_column* c = col_arg(args, 0);
    PyObject* o = PyTuple_GetItem(args, 1);
    custr_column* targets = nullptr;
    bool own = false;
    if (PyList_Check(o)) { targets = column_from_list(o); own = true; }
    else targets = (custr_column*)ptr_arg(args, 1);
    if (!targets) return fail_none();
    void* devptr = ptr_arg(args, 2);
    const uint32_t n = custr_size(c), m = custr_size(targets);
    PyObject* ret = nullptr;
    int rc = 0;
    if (devptr) {
        GIL_FREE(rc = custr_find_multiple(c, targets, (int32_t*)devptr, 1));
        ret = rc <= CUSTR_ERR_INVALID ? fail_none() : PyLong_FromVoidPtr(devptr);
    } else {
        std::vector<int32_t> host((size_t)n * m + 1);
        GIL_FREE(rc = custr_find_multiple(c, targets, host.data(), 0));
        if (rc <= CUSTR_ERR_INVALID) ret = fail_none();
        else {
            ret = PyList_New(n);
            for (uint32_t i = 0; i < n; ++i) {
                PyObject* row = PyList_New(m);
                for (uint32_t k = 0; k < m; ++k) PyList_SetItem(row, k, PyLong_FromLong(host[(size_t)i * m + k]));
                PyList_SetItem(ret, i, row);
            }
        }
    }
    if (own) custr_column_free(targets);
    return ret;
}

// ---- split family: column-major n_split / n_rsplit (cptr, delimiter|None, n) -> list of column handles
template <typename F>
static PyObject* columns_of(const custr_column* c, F call)
{
    std::vector<custr_column*> out(64, nullptr);
    int k = 0;
    Py_BEGIN_ALLOW_THREADS
    k = call(out.data(), (int32_t)out.size());
    if (k > (int)out.size()) {  // more columns than the first guess: release and ask again
        for (custr_column* x : out) custr_column_free(x);
        out.assign((size_t)k, nullptr);
        k = call(out.data(), (int32_t)out.size());
    }
    Py_END_ALLOW_THREADS
    if (k < 0) return fail_none();
    out.resize((size_t)k);
    return handle_list(out);
}
static PyObject* n_split(PyObject*, PyObject* args)
{
    const custr_column* c = col_arg(args, 0);
    const char* d = str_arg(args, 1);
    const int n = (int)int_arg(args, 2, -1);
    return columns_of(c, [&](custr_column** o, int32_t cap) { return custr_split(c, d, n, o, cap); });
}
static PyObject* n_rsplit(PyObject*, PyObject* args)
{
    const custr_column* c = col_arg(args, 0);
    const char* d = str_arg(args, 1);
    const int n = (int)int_arg(args, 2, -1);
    return columns_of(c, [&](custr_column** o, int32_t cap) { return custr_rsplit(c, d, n, o, cap); });
}
static PyObject* n_findall(PyObject*, PyObject* args)
{
    const custr_column* c = col_arg(args, 0);
    const char* p = str_arg(args, 1);
    return columns_of(c, [&](custr_column** o, int32_t cap) { return custr_findall(c, p, o, cap); });
}
static PyObject* n_extract(PyObject*, PyObject* args)
{
    const custr_column* c = col_arg(args, 0);
    const char* p = str_arg(args, 1);
    return columns_of(c, [&](custr_column** o, int32_t cap) { return custr_extract(c, p, o, cap); });
}
// row-major: one handle per row (0 for a null row), views over ONE flat token column (the reference allocates N objects,
// split.cu:171-190)
template <typename F>
static PyObject* records_of(const custr_column* c, F call)
{
    const uint32_t n = custr_size(c);
    std::vector<int32_t> row_off(n + 1, 0);
    custr_column* flat = nullptr;
    int rc = 0;
    GIL_FREE(rc = call(&flat, row_off.data()));
    if (rc < 0 || !flat) return fail_none();
    const std::vector<uint8_t> valid = valid_rows(c);
    std::vector<custr_column*> rows(n, nullptr);
    for (uint32_t i = 0; i < n; ++i)
        if (valid[i]) rows[i] = custr_slice_rows(flat, row_off[i], row_off[i + 1]);
    custr_column_free(flat);  // the views share its buffers
    return handle_list(rows);
}
static PyObject* n_split_record(PyObject*, PyObject* args)
{
    const custr_column* c = col_arg(args, 0);
    const char* d = str_arg(args, 1);
    const int n = (int)int_arg(args, 2, -1);
    return records_of(c, [&](custr_column** t, int32_t* ro) { return custr_split_record(c, d, n, t, ro, 0); });
}
static PyObject* n_rsplit_record(PyObject*, PyObject* args)
{
    const custr_column* c = col_arg(args, 0);
    const char* d = str_arg(args, 1);
    const int n = (int)int_arg(args, 2, -1);
    return records_of(c, [&](custr_column** t, int32_t* ro) { return custr_rsplit_record(c, d, n, t, ro, 0); });
}
static PyObject* n_findall_record(PyObject*, PyObject* args)
{
    const custr_column* c = col_arg(args, 0);
    const char* p = str_arg(args, 1);
    return records_of(c, [&](custr_column** t, int32_t* ro) { return custr_findall_record(c, p, t, ro, 0); });
}
// extract_record: per row one column of the groups of the first match = row i of every extract column
static PyObject* n_extract_record(PyObject*, PyObject* args)
{
    const custr_column* c = col_arg(args, 0);
    const char* p = str_arg(args, 1);
    std::vector<custr_column*> cols(64, nullptr);
    int k = 0;
    GIL_FREE(k = custr_extract(c, p, cols.data(), 64));
    if (k < 0) return fail_none();
    const uint32_t n = custr_size(c);
    const std::vector<uint8_t> valid = valid_rows(c);
    std::vector<custr_column*> rows(n, nullptr);
    // gather row i of each group column: build per-row columns from host strings (group counts are small)
    std::vector<PyObject*> lists;
    for (int g = 0; g < k && g < 64; ++g) lists.push_back(host_strings(cols[g]));
    for (uint32_t i = 0; i < n; ++i) {
        if (!valid[i]) continue;
        PyObject* row = PyList_New((Py_ssize_t)lists.size());
        for (size_t g = 0; g < lists.size(); ++g) {
            PyObject* v = PyList_GetItem(lists[g], i);
            Py_INCREF(v);
            PyList_SetItem(row, (Py_ssize_t)g, v);
        }
        rows[i] = column_from_list(row);
        Py_DECREF(row);
    }
    for (PyObject* l : lists) Py_DECREF(l);
    for (int g = 0; g < k && g < 64; ++g) custr_column_free(cols[g]);
    return handle_list(rows);
}
// n_partition / n_rpartition (cptr, delimiter) -> one 3-row handle per row (three nulls for a null row, like the reference)
static PyObject* partition_like(PyObject* args, int right)
{
    const custr_column* c = col_arg(args, 0);
    const char* d = str_arg(args, 1);
    custr_column* flat = nullptr;
    GIL_FREE(flat = custr_partition(c, d, right));
    if (!flat) return fail_none();
    const uint32_t n = custr_size(c);
    std::vector<custr_column*> rows(n, nullptr);
    for (uint32_t i = 0; i < n; ++i) rows[i] = custr_slice_rows(flat, (int32_t)(3 * i), (int32_t)(3 * i + 3));
    custr_column_free(flat);
    return handle_list(rows);
}
static PyObject* n_partition(PyObject*, PyObject* args) { return partition_like(args, 0); }
static PyObject* n_rpartition(PyObject*, PyObject* args) { return partition_like(args, 1); }

static PyMethodDef k_methods[] = {
    {"n_createFromHostStrings", n_createFromHostStrings, METH_VARARGS, ""}, {"n_destroyStrings", n_destroyStrings, METH_VARARGS, ""},
    {"n_createHostStrings", n_createHostStrings, METH_VARARGS, ""}, {"n_createFromOffsets", n_createFromOffsets, METH_VARARGS, ""},
    {"n_create_offsets", n_create_offsets, METH_VARARGS, ""}, {"n_size", n_size, METH_VARARGS, ""}, {"n_len", n_len, METH_VARARGS, ""},
    {"n_byte_count", n_byte_count, METH_VARARGS, ""}, {"n_set_null_bitmask", n_set_null_bitmask, METH_VARARGS, ""},
    {"n_null_count", n_null_count, METH_VARARGS, ""}, {"n_hash", n_hash, METH_VARARGS, ""}, {"n_copy", n_copy, METH_VARARGS, ""},
    {"n_gather", n_gather, METH_VARARGS, ""}, {"n_contains", n_contains, METH_VARARGS, ""}, {"n_match", n_match, METH_VARARGS, ""},
    {"n_count", n_count, METH_VARARGS, ""}, {"n_replace", n_replace, METH_VARARGS, ""}, {"n_replace_multi", n_replace_multi, METH_VARARGS, ""},
    {"n_replace_with_backrefs", n_replace_with_backrefs, METH_VARARGS, ""}, {"n_find", n_find, METH_VARARGS, ""},
    {"n_rfind", n_rfind, METH_VARARGS, ""}, {"n_find_from", n_find_from, METH_VARARGS, ""}, {"n_startswith", n_startswith, METH_VARARGS, ""},
    {"n_endswith", n_endswith, METH_VARARGS, ""}, {"n_match_strings", n_match_strings, METH_VARARGS, ""},
    {"n_find_multiple", n_find_multiple, METH_VARARGS, ""}, {"n_split", n_split, METH_VARARGS, ""}, {"n_rsplit", n_rsplit, METH_VARARGS, ""},
    {"n_split_record", n_split_record, METH_VARARGS, ""}, {"n_rsplit_record", n_rsplit_record, METH_VARARGS, ""},
    {"n_partition", n_partition, METH_VARARGS, ""}, {"n_rpartition", n_rpartition, METH_VARARGS, ""}, {"n_findall", n_findall, METH_VARARGS, ""},
    {"n_findall_record", n_findall_record, METH_VARARGS, ""}, {"n_extract", n_extract, METH_VARARGS, ""},
    {"n_extract_record", n_extract_record, METH_VARARGS, ""}, {nullptr, nullptr, 0, nullptr}};
static struct PyModuleDef k_module = {PyModuleDef_HEAD_INIT, "pyniNVStrings", "NVStrings hot path over libcustr.so (custrings_b200)", -1, k_methods};
PyMODINIT_FUNC PyInit_pyniNVStrings(void) { return PyModule_Create(&k_module); }

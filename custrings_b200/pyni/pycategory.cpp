// pyniNVCategory — hot-path subset of the reference's python/cpp/pycategory.cpp (method table :900-937) over libcustr.so.
#include "pyni_common.h"
using namespace pyni;
#define GIL_FREE(stmt) Py_BEGIN_ALLOW_THREADS stmt; Py_END_ALLOW_THREADS

// the reference recognises nvstrings instances by TYPE NAME and reads their m_cptr attribute (pycategory.cpp:42-71)
static bool strings_handles(PyObject* o, std::vector<const custr_column*>& out)
{
    auto one = [&](PyObject* x) -> bool {
        if (std::string(Py_TYPE(x)->tp_name) != "nvstrings") {
            PyErr_SetString(PyExc_ValueError, "nvcategory: argument must be nvstrings objects");
            return false;
        }
        PyObject* a = PyObject_GetAttrString(x, "m_cptr");
        void* p = a ? PyLong_AsVoidPtr(a) : nullptr;
        Py_XDECREF(a);
        if (!p) { PyErr_SetString(PyExc_ValueError, "nvcategory: invalid nvstrings object"); return false; }
        out.push_back((const custr_column*)p);
        return true;
    };
    if (!o || o == Py_None) { PyErr_SetString(PyExc_ValueError, "nvcategory: parameter required"); return false; }
    if (PyList_Check(o)) {
        for (Py_ssize_t i = 0; i < PyList_Size(o); ++i)
            if (!one(PyList_GetItem(o, i))) return false;
        return true;
    }
    return one(o);
}
static PyObject* n_createCategoryFromNVStrings(PyObject*, PyObject* args)
{
    std::vector<const custr_column*> cols;
    if (!strings_handles(PyTuple_GetItem(args, 0), cols)) return nullptr;
    custr_category* c = nullptr;
    GIL_FREE(c = custr_category_create(cols.data(), (int32_t)cols.size()));
    return handle_or_none(c);
}
static PyObject* n_createCategoryFromHostStrings(PyObject*, PyObject* args)
{
    custr_column* col = column_from_list(PyTuple_GetItem(args, 0));
    if (!col) return fail_none();
    const custr_column* cols[1] = {col};
    custr_category* c = nullptr;
    GIL_FREE(c = custr_category_create(cols, 1));
    custr_column_free(col);
    return handle_or_none(c);
}
static PyObject* n_destroyCategory(PyObject*, PyObject* args)
{
    custr_category* c = (custr_category*)ptr_arg(args, 0);
    GIL_FREE(custr_category_free(c));
    return PyLong_FromLong(0);
}
static PyObject* n_size(PyObject*, PyObject* args) { return PyLong_FromUnsignedLong(custr_category_size((const custr_category*)ptr_arg(args, 0))); }
static PyObject* n_keys_size(PyObject*, PyObject* args) { return PyLong_FromUnsignedLong(custr_category_keys_size((const custr_category*)ptr_arg(args, 0))); }
static PyObject* n_keys_type(PyObject*, PyObject*) { return PyUnicode_FromString("str"); }
static PyObject* n_get_keys(PyObject*, PyObject* args)
{
    custr_column* k = nullptr;
    const custr_category* c = (const custr_category*)ptr_arg(args, 0);
    GIL_FREE(k = custr_category_keys(c));
    return handle_or_none(k);
}
// (cptr, devptr): into the device array, or a host list
static PyObject* n_get_values(PyObject*, PyObject* args)
{
    const custr_category* c = (const custr_category*)ptr_arg(args, 0);
    void* devptr = ptr_arg(args, 1);
    int rc = 0;
    if (devptr) {
        GIL_FREE(rc = custr_category_values(c, (int32_t*)devptr, 1));
        if (rc <= CUSTR_ERR_INVALID) return fail_none();
        return PyLong_FromVoidPtr(devptr);
    }
    const uint32_t n = custr_category_size(c);
    std::vector<int32_t> v(n ? n : 1);
    GIL_FREE(rc = custr_category_values(c, v.data(), 0));
    if (rc <= CUSTR_ERR_INVALID) return fail_none();
    PyObject* list = PyList_New(n);
    for (uint32_t i = 0; i < n; ++i) PyList_SetItem(list, i, PyLong_FromLong(v[i]));
    return list;
}
static PyObject* n_get_values_cpointer(PyObject*, PyObject* args)
{
    return PyLong_FromVoidPtr((void*)custr_category_values_cptr((const custr_category*)ptr_arg(args, 0)));
}
// to_strings: keys gathered by the values (NVCategory.cu:977-1009)
static PyObject* n_to_strings(PyObject*, PyObject* args)
{
    const custr_category* c = (const custr_category*)ptr_arg(args, 0);
    custr_column *keys = nullptr, *out = nullptr;
    Py_BEGIN_ALLOW_THREADS
    keys = custr_category_keys(c);
    if (keys) out = custr_gather(keys, custr_category_values_cptr(c), (int32_t)custr_category_size(c), 1);
    custr_column_free(keys);
    Py_END_ALLOW_THREADS
    return handle_or_none(out);
}
static PyObject* merge_like(PyObject* args, int sorted)
{
    const custr_category* cats[2] = {(const custr_category*)ptr_arg(args, 0), nullptr};
    PyObject* other = PyTuple_GetItem(args, 1);
    if (PyLong_Check(other)) cats[1] = (const custr_category*)PyLong_AsVoidPtr(other);
    else {
        PyObject* a = PyObject_GetAttrString(other, "m_cptr");
        cats[1] = a ? (const custr_category*)PyLong_AsVoidPtr(a) : nullptr;
        Py_XDECREF(a);
    }
    custr_category* r = nullptr;
    GIL_FREE(r = custr_category_merge(cats, 2, sorted));
    return handle_or_none(r);
}
static PyObject* n_merge_category(PyObject*, PyObject* args) { return merge_like(args, 0); }
static PyObject* n_merge_and_remap(PyObject*, PyObject* args) { return merge_like(args, 1); }

static PyMethodDef k_methods[] = {
    {"n_createCategoryFromNVStrings", n_createCategoryFromNVStrings, METH_VARARGS, ""},
    {"n_createCategoryFromHostStrings", n_createCategoryFromHostStrings, METH_VARARGS, ""},
    {"n_destroyCategory", n_destroyCategory, METH_VARARGS, ""}, {"n_size", n_size, METH_VARARGS, ""},
    {"n_keys_size", n_keys_size, METH_VARARGS, ""}, {"n_keys_type", n_keys_type, METH_VARARGS, ""}, {"n_get_keys", n_get_keys, METH_VARARGS, ""},
    {"n_get_values", n_get_values, METH_VARARGS, ""}, {"n_get_values_cpointer", n_get_values_cpointer, METH_VARARGS, ""},
    {"n_to_strings", n_to_strings, METH_VARARGS, ""}, {"n_merge_category", n_merge_category, METH_VARARGS, ""},
    {"n_merge_and_remap", n_merge_and_remap, METH_VARARGS, ""}, {nullptr, nullptr, 0, nullptr}};
static struct PyModuleDef k_module = {PyModuleDef_HEAD_INIT, "pyniNVCategory", "NVCategory hot path over libcustr.so (custrings_b200)", -1, k_methods};
PyMODINIT_FUNC PyInit_pyniNVCategory(void) { return PyModule_Create(&k_module); }

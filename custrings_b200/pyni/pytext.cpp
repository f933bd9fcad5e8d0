// pyniNVText — hot-path subset of the reference's python/cpp/pytext.cpp (method table :653-674) over libcustr.so:
// n_tokenize (strs, delimiter|None), n_token_count (strs, delimiter|None, devptr).  `strs` is the nvstrings OBJECT (the
// reference reads its m_cptr attribute).
#include "pyni_common.h"
using namespace pyni;
#define GIL_FREE(stmt) Py_BEGIN_ALLOW_THREADS stmt; Py_END_ALLOW_THREADS

static const custr_column* strs_of(PyObject* o)
{
    if (!o || o == Py_None) return nullptr;
    if (PyLong_Check(o)) return (const custr_column*)PyLong_AsVoidPtr(o);
    PyObject* a = PyObject_GetAttrString(o, "m_cptr");
    const custr_column* c = a ? (const custr_column*)PyLong_AsVoidPtr(a) : nullptr;
    Py_XDECREF(a);
    return c;
}
static PyObject* n_tokenize(PyObject*, PyObject* args)
{
    const custr_column* c = strs_of(PyTuple_GetItem(args, 0));
    if (!c) { PyErr_SetString(PyExc_ValueError, "nvtext: invalid nvstrings object"); return nullptr; }
    const char* d = str_arg(args, 1);
    custr_column* r = nullptr;
    GIL_FREE(r = custr_tokenize(c, d));
    return handle_or_none(r);
}
static PyObject* n_token_count(PyObject*, PyObject* args)
{
    const custr_column* c = strs_of(PyTuple_GetItem(args, 0));
    if (!c) { PyErr_SetString(PyExc_ValueError, "nvtext: invalid nvstrings object"); return nullptr; }
    const char* d = str_arg(args, 1);
    return row_results<uint32_t>(c, ptr_arg(args, 2), 'u', 0, [&](uint32_t* o, int dm) { return (long long)custr_token_count(c, d, o, dm); });
}
static PyMethodDef k_methods[] = {{"n_tokenize", n_tokenize, METH_VARARGS, ""}, {"n_token_count", n_token_count, METH_VARARGS, ""},
                                  {nullptr, nullptr, 0, nullptr}};
static struct PyModuleDef k_module = {PyModuleDef_HEAD_INIT, "pyniNVText", "NVText hot path over libcustr.so (custrings_b200)", -1, k_methods};
PyMODINIT_FUNC PyInit_pyniNVText(void) { return PyModule_Create(&k_module); }

// seqident_ext.cpp -- the CPython extension `_seqident` of the B200 build (PyInit__seqident), the drop-in for the
// reference's sequali._seqident (_seqidentmodule.c:279-381): sequence_identity(target, query, match_score=1,
// mismatch_penalty=-1, deletion_penalty=-1, insertion_penalty=-1) -> float, on libsqgpu's k_seqident.
// Extension of this build: sequence_identities(pairs, ...) -> list of floats, any number of pairs in ONE launch.
#define PY_SSIZE_T_CLEAN
#include <Python.h>

#include <cmath>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/sqgpu.h"

namespace {

sq_ctx *g_ctx = nullptr;

sq_ctx *ctx_get() {
    if (g_ctx) return g_ctx;
    const int n = sq_device_count();
    if (n <= 0) {
        PyErr_SetString(PyExc_RuntimeError, "no CUDA device visible: sequali_b200 needs a GPU (there is no CPU fallback)");
        return nullptr;
    }
    const char *dev = getenv("SEQUALI_B200_DEVICE");
    if (!dev) dev = getenv("LOCAL_RANK");
    int device = dev ? atoi(dev) : 0;
    if (device < 0) device = 0;
    if (sq_ctx_create(device % n, &g_ctx) != SQ_OK) {
        g_ctx = nullptr;
        PyErr_Format(PyExc_RuntimeError, "sq_ctx_create: %s", sq_last_error());
        return nullptr;
    }
    return g_ctx;
}

struct Pairs {
    std::string targets, queries;
    std::vector<uint64_t> t_off{0};
    std::vector<uint32_t> q_off{0};
};

// appends one pair; the checks and messages of _seqidentmodule.c:311-337
int add_pair(Pairs &p, PyObject *target_obj, PyObject *query_obj) {
    Py_ssize_t t_utf8 = 0, q_utf8 = 0;
    const char *t = PyUnicode_AsUTF8AndSize(target_obj, &t_utf8);
    if (!t) return -1;
    const char *q = PyUnicode_AsUTF8AndSize(query_obj, &q_utf8);
    if (!q) return -1;
    if (PyUnicode_GetLength(target_obj) != t_utf8 || PyUnicode_GetLength(query_obj) != q_utf8) {
        PyErr_Format(PyExc_ValueError, "Only ascii strings are allowed. Got %R", target_obj);
        return -1;
    }
    if (q_utf8 > 31) {
        PyErr_Format(PyExc_ValueError, "Only query with lengths less than 32 are supported. Got %zd", q_utf8);
        return -1;
    }
    p.targets.append(t, (size_t)t_utf8);
    p.queries.append(q, (size_t)q_utf8);
    p.t_off.push_back(p.targets.size());
    p.q_off.push_back((uint32_t)p.queries.size());
    return 0;
}

int run(const Pairs &p, Py_ssize_t match, Py_ssize_t mismatch, Py_ssize_t deletion, Py_ssize_t insertion,
        std::vector<int32_t> &out) {
    sq_ctx *ctx = ctx_get();
    if (!ctx) return -1;
    const uint64_t n = p.q_off.size() - 1;
    out.assign(n, 0);
    int rc;
    Py_BEGIN_ALLOW_THREADS
    // the reference hands the scores on as int8_t
    rc = sq_sequence_identity_batch(ctx, (const uint8_t *)p.targets.data(), p.t_off.data(), (const uint8_t *)p.queries.data(),
                                    p.q_off.data(), n, (int8_t)match, (int8_t)mismatch, (int8_t)deletion, (int8_t)insertion,
                                    out.data());
    Py_END_ALLOW_THREADS
    if (rc != SQ_OK) {
        PyErr_Format(PyExc_RuntimeError, "sq_sequence_identity_batch: %s", sq_last_error());
        return -1;
    }
    return 0;
}

double identity(int32_t matches, uint32_t query_length) {
    return query_length ? (double)matches / (double)query_length : std::nan("");  // 0 / 0 in the reference
}

const char *kwnames[] = {"target", "query", "match_score", "mismatch_penalty", "deletion_penalty", "insertion_penalty", nullptr};

PyObject *sequence_identity(PyObject *, PyObject *args, PyObject *kwargs) {
    PyObject *target_obj = nullptr, *query_obj = nullptr;
    Py_ssize_t match = 1, mismatch = -1, deletion = -1, insertion = -1;
    if (!PyArg_ParseTupleAndKeywords(args, kwargs, "UU|nnnn:identify_sequence", (char **)kwnames, &target_obj, &query_obj,
                                     &match, &mismatch, &deletion, &insertion))
        return nullptr;
    Pairs p;
    std::vector<int32_t> out;
    if (add_pair(p, target_obj, query_obj) < 0 || run(p, match, mismatch, deletion, insertion, out) < 0) return nullptr;
    return PyFloat_FromDouble(identity(out[0], p.q_off[1]));
}

const char *batch_kwnames[] = {"pairs", "match_score", "mismatch_penalty", "deletion_penalty", "insertion_penalty", nullptr};

PyObject *sequence_identities(PyObject *, PyObject *args, PyObject *kwargs) {
    PyObject *pairs_obj = nullptr;
    Py_ssize_t match = 1, mismatch = -1, deletion = -1, insertion = -1;
    if (!PyArg_ParseTupleAndKeywords(args, kwargs, "O|nnnn:sequence_identities", (char **)batch_kwnames, &pairs_obj, &match,
                                     &mismatch, &deletion, &insertion))
        return nullptr;
    PyObject *it = PyObject_GetIter(pairs_obj);
    if (!it) return nullptr;
    Pairs p;
    while (PyObject *item = PyIter_Next(it)) {
        PyObject *t = nullptr, *q = nullptr;
        int ok = PyArg_ParseTuple(item, "UU:sequence_identities", &t, &q) && add_pair(p, t, q) == 0;
        Py_DECREF(item);
        if (!ok) {
            Py_DECREF(it);
            return nullptr;
        }
    }
    Py_DECREF(it);
    if (PyErr_Occurred()) return nullptr;
    const size_t n = p.q_off.size() - 1;
    std::vector<int32_t> out;
    if (n && run(p, match, mismatch, deletion, insertion, out) < 0) return nullptr;
    PyObject *list = PyList_New((Py_ssize_t)n);
    if (!list) return nullptr;
    for (size_t i = 0; i < n; i++) {
        PyObject *f = PyFloat_FromDouble(identity(out[i], p.q_off[i + 1] - p.q_off[i]));
        if (!f) {
            Py_DECREF(list);
            return nullptr;
        }
        PyList_SET_ITEM(list, (Py_ssize_t)i, f);
    }
    return list;
}

PyMethodDef methods[] = {
    {"sequence_identity", (PyCFunction)(void (*)(void))sequence_identity, METH_VARARGS | METH_KEYWORDS,
     "Calculate sequence identity based on a smith-waterman matrix (on the device).\n"
     "Identity is given as (query_length - errors / query_length).\n"},
    {"sequence_identities", (PyCFunction)(void (*)(void))sequence_identities, METH_VARARGS | METH_KEYWORDS,
     "sequence_identity for an iterable of (target, query) pairs in one launch (extension of the B200 build)."},
    {nullptr, nullptr, 0, nullptr}};

PyModuleDef_Slot slots[] = {{0, nullptr}};
PyModuleDef module = {PyModuleDef_HEAD_INIT, "_seqident", nullptr, 0, methods, slots, nullptr, nullptr, nullptr};

}  // namespace

extern "C" __attribute__((visibility("default"))) PyObject *PyInit__seqident(void) { return PyModuleDef_Init(&module); }

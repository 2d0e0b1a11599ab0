// qc_ext.cpp -- the CPython extension `_qc` of the B200 build.
//
// The reference's only native module is `sequali._qc` (src/sequali/_qcmodule.c:5984-6234: thirteen
// heap types + module constants; stubs src/sequali/_qc.pyi:45-189).  This file is its counterpart on
// top of libsqgpu.so (include/sqgpu.h): the same type names, constructor arguments, methods,
// members, constants and exception types / messages, with the per-record work done by the sm_100a
// kernels behind the C ABI.  It holds no algorithm of the hot path: it moves bytes between Python
// objects and the ABI and turns status structs into the reference's exceptions.
//
//   * `add_record_array[_pair]` is deferred (results are observable only through
//     getters and members): the collectors fed with one record array are gathered and handed to
//     the device in ONE sq_fused_add call.  Getters, members (PyGetSetDef that flush first),
//     `add_read`, or another record array flush.
//   * `add_read` / `add_sequence[_pair]` are synchronous: tests expect warnings and errors inside
//     the call.
//   * no CPU fallback: without a CUDA device every constructor raises SqGpuError.
//
// Built by sequali_b200/ext/Makefile as `_qc.so` (PyInit__qc): importable as `sequali_b200.ext._qc`
// and, linked into a directory named `sequali` next to the reference's unchanged __init__.py, as
// `sequali._qc`.
#define PY_SSIZE_T_CLEAN
#include <Python.h>

#include <cerrno>
#include <sys/stat.h>
#include <sys/types.h>
#include <unistd.h>
#include <structmember.h>

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sqgpu.h"

namespace {

// ---------------------------------------------------------------------------------------------
// context, errors
// ---------------------------------------------------------------------------------------------
PyObject *SqGpuError = nullptr;
PyObject *g_array_type = nullptr;  // array.array
sq_ctx *g_ctx = nullptr;

// status code of libsqgpu -> Python exception (the ctypes mirror's `check`)
PyObject *raise_sq(int rc, const char *what) {
    const char *msg = sq_last_error();
    if (rc == SQ_E_NOMEM) PyErr_Format(PyExc_MemoryError, "%s: %s", what, msg);
    else if (rc == SQ_E_ARG) PyErr_Format(PyExc_ValueError, "%s: %s", what, msg);
    else PyErr_Format(SqGpuError, "%s failed (code %d): %s", what, rc, msg);
    return nullptr;
}
#define SQ_CHECK(expr, what)                  \
    do {                                      \
        int _rc = (expr);                     \
        if (_rc != SQ_OK) return raise_sq(_rc, what); \
    } while (0)
#define SQ_CHECK_INT(expr, what)              \
    do {                                      \
        int _rc = (expr);                     \
        if (_rc != SQ_OK) {                   \
            raise_sq(_rc, what);              \
            return -1;                        \
        }                                     \
    } while (0)

sq_ctx *ctx_get() {
    if (g_ctx) return g_ctx;
    const int n = sq_device_count();
    if (n <= 0) {
        PyErr_SetString(SqGpuError, "no CUDA device visible: sequali_b200 needs a GPU (there is no CPU fallback)");
        return nullptr;
    }
    const char *dev = getenv("SEQUALI_B200_DEVICE");
    if (!dev) dev = getenv("LOCAL_RANK");
    int device = dev ? atoi(dev) : 0;
    if (device < 0) device = 0;
    int rc = sq_ctx_create(device % n, &g_ctx);
    if (rc != SQ_OK) {
        g_ctx = nullptr;
        raise_sq(rc, "sq_ctx_create");
        return nullptr;
    }
    return g_ctx;
}

// ---- pinned staging buffers: size classes, one global byte cap, least recently used out first ----
struct PinnedPool {
    struct Item { void *ptr; size_t size; };
    std::vector<Item> free_list;  // oldest first
    size_t bytes = 0;
    static constexpr size_t CAP = (size_t)2 << 30;  // idle bytes kept; the oldest blocks go first
    static size_t size_class(size_t n) {
        if (n < 4096) n = 4096;
        if (n >= ((size_t)1 << 20)) return (n + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
        size_t c = 4096;
        while (c < n) c <<= 1;
        return c;
    }
    void *take(size_t cls) {
        for (size_t i = free_list.size(); i-- > 0;)
            if (free_list[i].size == cls) {
                void *p = free_list[i].ptr;
                free_list.erase(free_list.begin() + i);
                bytes -= cls;
                return p;
            }
        return nullptr;
    }
    void give(void *p, size_t cls) {
        free_list.push_back({p, cls});
        bytes += cls;
        while (bytes > CAP && !free_list.empty()) {
            sq_pinned_free(g_ctx, free_list.front().ptr);
            bytes -= free_list.front().size;
            free_list.erase(free_list.begin());
        }
    }
} g_pool;

struct Pinned {
    uint8_t *ptr = nullptr;
    size_t size = 0, cls = 0;
    bool alloc(size_t n) {
        cls = PinnedPool::size_class(n);
        ptr = (uint8_t *)g_pool.take(cls);
        if (!ptr) ptr = (uint8_t *)sq_pinned_alloc(g_ctx, cls);
        if (!ptr) {
            PyErr_SetString(PyExc_MemoryError, sq_last_error());
            return false;
        }
        size = n;
        return true;
    }
    void release() {
        if (ptr) g_pool.give(ptr, cls);
        ptr = nullptr;
        size = cls = 0;
    }
};

PyObject *u64_array(const uint64_t *data, size_t n) {
    PyObject *arr = PyObject_CallFunction(g_array_type, "s", "Q");
    if (!arr) return nullptr;
    if (n) {
        PyObject *bytes = PyBytes_FromStringAndSize((const char *)data, (Py_ssize_t)(n * 8));
        if (!bytes) {
            Py_DECREF(arr);
            return nullptr;
        }
        PyObject *r = PyObject_CallMethod(arr, "frombytes", "O", bytes);
        Py_DECREF(bytes);
        if (!r) {
            Py_DECREF(arr);
            return nullptr;
        }
        Py_DECREF(r);
    }
    return arr;
}

// list of Python ints / floats from a plain array
PyObject *u64_list(const uint64_t *v, size_t n) {
    PyObject *l = PyList_New((Py_ssize_t)n);
    for (size_t i = 0; l && i < n; i++) {
        PyObject *o = PyLong_FromUnsignedLongLong(v[i]);
        if (!o) {
            Py_CLEAR(l);
            break;
        }
        PyList_SET_ITEM(l, (Py_ssize_t)i, o);
    }
    return l;
}
// dict from "key", value, ... (values are stolen; a NULL value fails the whole dict)
PyObject *dict_of(std::initializer_list<std::pair<const char *, PyObject *>> items) {
    PyObject *d = PyDict_New();
    bool ok = d != nullptr;
    for (auto &kv : items) {
        if (ok && (!kv.second || PyDict_SetItemString(d, kv.first, kv.second) < 0)) ok = false;
        Py_XDECREF(kv.second);
    }
    if (!ok) Py_CLEAR(d);
    return d;
}
// uint64 values of an iterable of ints (flattening one level of pairs when `pairs`)
int u64_from_iterable(PyObject *obj, bool pairs, std::vector<uint64_t> &out) {
    PyObject *it = PyObject_GetIter(obj);
    if (!it) return -1;
    while (PyObject *item = PyIter_Next(it)) {
        if (pairs) {
            unsigned long long a, b;
            if (!PyArg_ParseTuple(item, "KK", &a, &b)) {
                Py_DECREF(item);
                Py_DECREF(it);
                return -1;
            }
            out.push_back(a);
            out.push_back(b);
        }
        else {
            const unsigned long long a = PyLong_AsUnsignedLongLong(item);
            if (a == (unsigned long long)-1 && PyErr_Occurred()) {
                Py_DECREF(item);
                Py_DECREF(it);
                return -1;
            }
            out.push_back(a);
        }
        Py_DECREF(item);
    }
    Py_DECREF(it);
    return PyErr_Occurred() ? -1 : 0;
}

const double *error_rates() {
    static double tab[94];
    static bool done = false;
    if (!done) {
        for (int q = 0; q < 94; q++) tab[q] = pow(10.0, -((double)q / 10.0));
        done = true;
    }
    return tab;
}

// ---------------------------------------------------------------------------------------------
// FastqRecordView (reference :357-569)
// ---------------------------------------------------------------------------------------------
struct RecordView {
    PyObject_HEAD
    PyObject *obj;  // bytes
    uint32_t name_off, name_len, seq_off, seq_len, qual_off, tags_off, tags_len;
    double err;
};
PyTypeObject *RecordViewType = nullptr, *ArrayViewType = nullptr;

const char *ascii_of(PyObject *s, const char *label, Py_ssize_t *len, PyObject *for_msg) {
    if (!PyUnicode_Check(s)) {
        PyErr_Format(PyExc_TypeError, "FastqRecordView() argument '%s' must be str, not %s", label, Py_TYPE(s)->tp_name);
        return nullptr;
    }
    if (!PyUnicode_IS_COMPACT_ASCII(s)) {
        PyErr_Format(PyExc_ValueError, "%s should contain only ASCII characters: %R", label, for_msg);
        return nullptr;
    }
    return PyUnicode_AsUTF8AndSize(s, len);
}

PyObject *RecordView_new(PyTypeObject *type, PyObject *args, PyObject *kwargs) {
    static const char *kw[] = {"name", "sequence", "qualities", "tags", nullptr};
    PyObject *name = nullptr, *seq = nullptr, *qual = nullptr, *tags = Py_None;
    if (!PyArg_ParseTupleAndKeywords(args, kwargs, "OOO|O:FastqRecordView", (char **)kw, &name, &seq, &qual, &tags))
        return nullptr;
    Py_ssize_t ln, ls, lq, lt = 0;
    const char *pn = ascii_of(name, "name", &ln, name);
    if (!pn) return nullptr;
    const char *ps = ascii_of(seq, "sequence", &ls, seq);
    if (!ps) return nullptr;
    const char *pq = ascii_of(qual, "qualities", &lq, seq);  // (the reference prints the sequence here too)
    if (!pq) return nullptr;
    const char *pt = nullptr;
    if (tags != Py_None) {
        if (!PyBytes_Check(tags)) {
            PyErr_Format(PyExc_TypeError, "FastqRecordView() argument 'tags' must be bytes, not %s", Py_TYPE(tags)->tp_name);
            return nullptr;
        }
        pt = PyBytes_AS_STRING(tags);
        lt = PyBytes_GET_SIZE(tags);
    }
    if (ls != lq) {
        PyErr_Format(PyExc_ValueError, "sequence and qualities have different lengths: %zd and %zd", ls, lq);
        return nullptr;
    }
    const uint64_t total = (uint64_t)ln + 2 * (uint64_t)ls + (uint64_t)lt;
    if (total > 0xFFFFFFFFULL) {
        PyErr_Format(PyExc_OverflowError, "Total length of FASTQ record exceeds 4 GiB. Record name: %R", name);
        return nullptr;
    }
    double err = 0.0;  // eager validation + plain left-to-right sum (:442-451)
    const double *rates = error_rates();
    for (Py_ssize_t i = 0; i < lq; i++) {
        const unsigned q = (unsigned char)pq[i] - 33u;
        if (q > 93u) {
            PyErr_Format(PyExc_ValueError, "Not a valid phred character: %c", pq[i]);
            return nullptr;
        }
        err += rates[q];
    }
    PyObject *obj = PyBytes_FromStringAndSize(nullptr, (Py_ssize_t)total);
    if (!obj) return nullptr;
    char *o = PyBytes_AS_STRING(obj);
    memcpy(o, pn, ln);
    memcpy(o + ln, ps, ls);
    memcpy(o + ln + ls, pq, lq);
    if (lt) memcpy(o + ln + 2 * ls, pt, lt);
    RecordView *self = (RecordView *)type->tp_alloc(type, 0);
    if (!self) {
        Py_DECREF(obj);
        return nullptr;
    }
    self->obj = obj;
    self->name_off = 0;
    self->name_len = (uint32_t)ln;
    self->seq_off = (uint32_t)ln;
    self->seq_len = (uint32_t)ls;
    self->qual_off = (uint32_t)(ln + ls);
    self->tags_off = (uint32_t)(ln + 2 * ls);
    self->tags_len = (uint32_t)lt;
    self->err = err;
    return (PyObject *)self;
}
void RecordView_dealloc(RecordView *self) {
    PyTypeObject *tp = Py_TYPE(self);
    Py_XDECREF(self->obj);
    tp->tp_free((PyObject *)self);
    Py_DECREF(tp);
}
PyObject *RecordView_from_meta(PyObject *obj, const sq_meta &m) {
    RecordView *v = (RecordView *)RecordViewType->tp_alloc(RecordViewType, 0);
    if (!v) return nullptr;
    Py_INCREF(obj);
    v->obj = obj;
    v->name_off = m.name_off;
    v->name_len = m.name_len;
    v->seq_off = m.seq_off;
    v->seq_len = m.seq_len;
    v->qual_off = m.qual_off;
    v->tags_off = m.tags_off;
    v->tags_len = m.tags_len;
    v->err = m.err_sum;
    return (PyObject *)v;
}
PyObject *RecordView_name(RecordView *s, PyObject *) {
    return PyUnicode_DecodeASCII(PyBytes_AS_STRING(s->obj) + s->name_off, s->name_len, nullptr);
}
PyObject *RecordView_sequence(RecordView *s, PyObject *) {
    return PyUnicode_DecodeASCII(PyBytes_AS_STRING(s->obj) + s->seq_off, s->seq_len, nullptr);
}
PyObject *RecordView_qualities(RecordView *s, PyObject *) {
    return PyUnicode_DecodeASCII(PyBytes_AS_STRING(s->obj) + s->qual_off, s->seq_len, nullptr);
}
PyObject *RecordView_tags(RecordView *s, PyObject *) {
    return PyBytes_FromStringAndSize(PyBytes_AS_STRING(s->obj) + s->tags_off, s->tags_len);
}
PyMethodDef RecordView_methods[] = {
    {"name", (PyCFunction)RecordView_name, METH_NOARGS, "name of the record"},
    {"sequence", (PyCFunction)RecordView_sequence, METH_NOARGS, "sequence of the record"},
    {"qualities", (PyCFunction)RecordView_qualities, METH_NOARGS, "qualities of the record"},
    {"tags", (PyCFunction)RecordView_tags, METH_NOARGS, "raw BAM tags of the record"},
    {nullptr, nullptr, 0, nullptr}};
PyMemberDef RecordView_members[] = {
    {"obj", T_OBJECT, offsetof(RecordView, obj), READONLY, "the bytes object the record lives in"},
    {nullptr, 0, 0, 0, nullptr}};
PyType_Slot RecordView_slots[] = {{Py_tp_new, (void *)RecordView_new},
                                  {Py_tp_dealloc, (void *)RecordView_dealloc},
                                  {Py_tp_methods, RecordView_methods},
                                  {Py_tp_members, RecordView_members},
                                  {0, nullptr}};
PyType_Spec RecordView_spec = {"_qc.FastqRecordView", sizeof(RecordView), 0, Py_TPFLAGS_DEFAULT, RecordView_slots};

// ---------------------------------------------------------------------------------------------
// FastqRecordArrayView (reference :575-883) + the handle of its device-resident copy
// ---------------------------------------------------------------------------------------------
struct ArrayView {
    PyObject_HEAD
    PyObject *obj;      // bytes, or NULL while the text only lives in `pinned` / on the device
    sq_meta *metas;     // host descriptors, or NULL (still on the device)
    sq_batch *h;        // device record array, or NULL (not uploaded yet)
    uint64_t n, nbytes;
    Pinned pinned;      // staging buffer a parser-made array was read into
    size_t pinned_off;  // where the text starts in it
    bool metas_stale;   // err_sum changed on the device (QCMetrics ran)
};

int flush_pending();
struct PendingAdds {
    PyObject *array = nullptr;  // the record array being gathered (strong reference)
    PyObject *mods[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // qc pt ov ns ad dd
} g_pending;
enum { ROLE_QC = 0, ROLE_PT, ROLE_OV, ROLE_NS, ROLE_AD, ROLE_DD };

ArrayView *ArrayView_alloc() {
    ArrayView *a = (ArrayView *)ArrayViewType->tp_alloc(ArrayViewType, 0);
    if (!a) return nullptr;
    a->obj = nullptr;
    a->metas = nullptr;
    a->h = nullptr;
    a->n = a->nbytes = 0;
    new (&a->pinned) Pinned();
    a->pinned_off = 0;
    a->metas_stale = false;
    return a;
}
PyObject *ArrayView_empty() {
    ArrayView *a = ArrayView_alloc();
    if (!a) return nullptr;
    a->obj = PyBytes_FromStringAndSize("", 0);
    return (PyObject *)a;
}
void ArrayView_dealloc(ArrayView *self) {
    PyTypeObject *tp = Py_TYPE(self);
    if (self->h) sq_batch_free(self->h);
    self->pinned.release();
    free(self->metas);
    Py_XDECREF(self->obj);
    tp->tp_free((PyObject *)self);
    Py_DECREF(tp);
}
PyObject *ArrayView_new(PyTypeObject *type, PyObject *args, PyObject *kwargs) {
    static const char *kw[] = {"view_items", nullptr};
    PyObject *items = nullptr;
    if (!PyArg_ParseTupleAndKeywords(args, kwargs, "O:FastqRecordArrayView", (char **)kw, &items)) return nullptr;
    PyObject *seq = PySequence_Fast(items, "view_items should be iterable");
    if (!seq) return nullptr;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(seq);
    uint64_t total = 0;
    for (Py_ssize_t i = 0; i < n; i++) {
        PyObject *it = PySequence_Fast_GET_ITEM(seq, i);
        if (Py_TYPE(it) != RecordViewType) {
            PyErr_Format(PyExc_TypeError, "Expected an iterable of FastqRecordView objects, but item %zd is of type %R: %R",
                         i, (PyObject *)Py_TYPE(it), it);
            Py_DECREF(seq);
            return nullptr;
        }
        RecordView *v = (RecordView *)it;
        total += (uint64_t)v->name_len + 2 * (uint64_t)v->seq_len + v->tags_len;
    }
    ArrayView *a = ArrayView_alloc();
    PyObject *obj = a ? PyBytes_FromStringAndSize(nullptr, (Py_ssize_t)total) : nullptr;
    sq_meta *metas = obj ? (sq_meta *)calloc(n ? n : 1, sizeof(sq_meta)) : nullptr;
    if (!a || !obj || !metas) {
        Py_XDECREF((PyObject *)a);
        Py_XDECREF(obj);
        free(metas);
        Py_DECREF(seq);
        return a && obj ? PyErr_NoMemory() : nullptr;
    }
    char *o = PyBytes_AS_STRING(obj);
    uint64_t off = 0;
    for (Py_ssize_t i = 0; i < n; i++) {
        RecordView *v = (RecordView *)PySequence_Fast_GET_ITEM(seq, i);
        const char *src = PyBytes_AS_STRING(v->obj);
        sq_meta &m = metas[i];
        m.name_off = (uint32_t)off;
        m.name_len = v->name_len;
        memcpy(o + off, src + v->name_off, v->name_len);
        off += v->name_len;
        m.seq_off = (uint32_t)off;
        m.seq_len = v->seq_len;
        memcpy(o + off, src + v->seq_off, v->seq_len);
        off += v->seq_len;
        m.qual_off = (uint32_t)off;
        memcpy(o + off, src + v->qual_off, v->seq_len);
        off += v->seq_len;
        m.tags_off = (uint32_t)off;
        m.tags_len = v->tags_len;
        memcpy(o + off, src + v->tags_off, v->tags_len);
        off += v->tags_len;
        m.err_sum = v->err;
    }
    Py_DECREF(seq);
    a->obj = obj;
    a->metas = metas;
    a->n = (uint64_t)n;
    a->nbytes = total;
    (void)type;
    return (PyObject *)a;
}
// sq_batch* of the array, uploading a Python-built one on first use
sq_batch *ArrayView_handle(ArrayView *a) {
    if (a->h) return a->h;
    sq_ctx *ctx = ctx_get();
    if (!ctx) return nullptr;
    int rc = sq_batch_from_packed(ctx, (const uint8_t *)PyBytes_AS_STRING(a->obj), a->nbytes, a->metas, a->n, &a->h);
    if (rc != SQ_OK) {
        a->h = nullptr;
        raise_sq(rc, "sq_batch_from_packed");
        return nullptr;
    }
    return a->h;
}
int ArrayView_fetch_metas(ArrayView *a) {
    if (g_pending.array == (PyObject *)a && flush_pending() < 0) return -1;
    if (a->metas && !a->metas_stale) return 0;
    if (!a->metas) a->metas = (sq_meta *)calloc(a->n ? a->n : 1, sizeof(sq_meta));
    if (!a->metas) {
        PyErr_NoMemory();
        return -1;
    }
    if (a->n && a->h) SQ_CHECK_INT(sq_batch_get_metas(a->h, a->metas), "sq_batch_get_metas");
    a->metas_stale = false;
    return 0;
}
PyObject *ArrayView_obj(ArrayView *a, void *) {
    if (!a->obj) {
        if (a->pinned.ptr) a->obj = PyBytes_FromStringAndSize((const char *)a->pinned.ptr + a->pinned_off, (Py_ssize_t)a->nbytes);
        else {
            a->obj = PyBytes_FromStringAndSize(nullptr, (Py_ssize_t)a->nbytes);
            if (a->obj && a->nbytes) {
                int rc = sq_batch_get_bytes(a->h, (uint8_t *)PyBytes_AS_STRING(a->obj));
                if (rc != SQ_OK) {
                    Py_CLEAR(a->obj);
                    return raise_sq(rc, "sq_batch_get_bytes");
                }
            }
        }
        if (!a->obj) return nullptr;
    }
    Py_INCREF(a->obj);
    return a->obj;
}
Py_ssize_t ArrayView_len(ArrayView *a) { return (Py_ssize_t)a->n; }
PyObject *ArrayView_item(ArrayView *a, Py_ssize_t i) {
    if (i < 0 || (uint64_t)i >= a->n) {
        PyErr_SetString(PyExc_IndexError, "array index out of range");
        return nullptr;
    }
    PyObject *obj = ArrayView_obj(a, nullptr);
    if (!obj) return nullptr;
    PyObject *res = ArrayView_fetch_metas(a) < 0 ? nullptr : RecordView_from_meta(obj, a->metas[i]);
    Py_DECREF(obj);
    return res;
}
PyObject *ArrayView_subscript(ArrayView *a, PyObject *key) {
    Py_ssize_t i = PyNumber_AsSsize_t(key, PyExc_IndexError);
    if (i == -1 && PyErr_Occurred()) return nullptr;
    if (i < 0) i += (Py_ssize_t)a->n;
    return ArrayView_item(a, i);
}
PyObject *ArrayView_is_mate(ArrayView *a, PyObject *other_o) {
    if (Py_TYPE(other_o) != ArrayViewType) {
        PyErr_Format(PyExc_TypeError, "other must be of type FastqRecordArrayView, got %R", (PyObject *)Py_TYPE(other_o));
        return nullptr;
    }
    ArrayView *b = (ArrayView *)other_o;
    if (a->n != b->n) {
        PyErr_Format(PyExc_ValueError,
                     "other is not the same length as this record array view. This length: %zd, other length: %zd",
                     (Py_ssize_t)a->n, (Py_ssize_t)b->n);
        return nullptr;
    }
    if (a->n == 0) Py_RETURN_TRUE;
    if (flush_pending() < 0) return nullptr;
    sq_batch *ha = ArrayView_handle(a), *hb = ha ? ArrayView_handle(b) : nullptr;
    if (!hb) return nullptr;
    uint64_t first = 0;
    int rc;
    Py_BEGIN_ALLOW_THREADS
    rc = sq_batch_is_mate(ha, hb, &first);
    Py_END_ALLOW_THREADS
    SQ_CHECK(rc, "sq_batch_is_mate");
    return PyBool_FromLong(first == a->n);
}
PyMethodDef ArrayView_methods[] = {
    {"is_mate", (PyCFunction)ArrayView_is_mate, METH_O, "Check if the record IDs in this array match those of other"},
    {nullptr, nullptr, 0, nullptr}};
PyGetSetDef ArrayView_getset[] = {{"obj", (getter)ArrayView_obj, nullptr, "the bytes object holding the records", nullptr},
                                  {nullptr, nullptr, nullptr, nullptr, nullptr}};
PyType_Slot ArrayView_slots[] = {{Py_tp_new, (void *)ArrayView_new},
                                 {Py_tp_dealloc, (void *)ArrayView_dealloc},
                                 {Py_tp_methods, ArrayView_methods},
                                 {Py_tp_getset, ArrayView_getset},
                                 {Py_sq_length, (void *)ArrayView_len},
                                 {Py_sq_item, (void *)ArrayView_item},
                                 {Py_mp_length, (void *)ArrayView_len},
                                 {Py_mp_subscript, (void *)ArrayView_subscript},
                                 {0, nullptr}};
PyType_Spec ArrayView_spec = {"_qc.FastqRecordArrayView", sizeof(ArrayView), 0, Py_TPFLAGS_DEFAULT, ArrayView_slots};

ArrayView *check_array(PyObject *o, const char *name) {
    if (Py_TYPE(o) != ArrayViewType) {
        PyErr_Format(PyExc_TypeError, "%s should be a FastqRecordArrayView object, got %R", name, (PyObject *)Py_TYPE(o));
        return nullptr;
    }
    return (ArrayView *)o;
}
RecordView *check_read(PyObject *o) {
    if (Py_TYPE(o) != RecordViewType) {
        PyErr_Format(PyExc_TypeError, "read should be a FastqRecordView object, got %R", (PyObject *)Py_TYPE(o));
        return nullptr;
    }
    return (RecordView *)o;
}
// a one-record array around `read` (add_read = stage + synchronous flush)
ArrayView *single_array(RecordView *v) {
    PyObject *tup = PyTuple_Pack(1, (PyObject *)v);
    if (!tup) return nullptr;
    PyObject *args = PyTuple_Pack(1, tup);
    Py_DECREF(tup);
    if (!args) return nullptr;
    PyObject *a = ArrayView_new(ArrayViewType, args, nullptr);
    Py_DECREF(args);
    return (ArrayView *)a;
}
// a one-record array holding only a sequence (DedupEstimator / InsertSizeMetrics add_sequence*)
ArrayView *sequence_array(PyObject *s, const char *what) {
    if (!PyUnicode_Check(s)) {
        PyErr_Format(PyExc_TypeError, "sequence should be a str object, got %R", (PyObject *)Py_TYPE(s));
        return nullptr;
    }
    if (!PyUnicode_IS_COMPACT_ASCII(s)) {
        PyErr_Format(PyExc_ValueError, "%s should consist only of ASCII characters.", what);
        return nullptr;
    }
    Py_ssize_t len;
    const char *p = PyUnicode_AsUTF8AndSize(s, &len);
    ArrayView *a = ArrayView_alloc();
    if (!a) return nullptr;
    a->obj = PyBytes_FromStringAndSize(p, len);
    a->metas = (sq_meta *)calloc(1, sizeof(sq_meta));
    if (!a->obj || !a->metas) {
        Py_DECREF((PyObject *)a);
        return (ArrayView *)PyErr_NoMemory();
    }
    a->metas[0].seq_len = (uint32_t)len;
    a->metas[0].tags_off = (uint32_t)len;
    a->n = 1;
    a->nbytes = (uint64_t)len;
    return a;
}

// ---------------------------------------------------------------------------------------------
// collectors: common head, deferred adds
// ---------------------------------------------------------------------------------------------
struct Collector {
    PyObject_HEAD
    void *h;
};

int flush_pending() {
    if (!g_pending.array) return 0;
    ArrayView *arr = (ArrayView *)g_pending.array;
    PyObject *mods[6];
    for (int r = 0; r < 6; r++) {
        mods[r] = g_pending.mods[r];
        g_pending.mods[r] = nullptr;
    }
    g_pending.array = nullptr;
    sq_batch *b = ArrayView_handle(arr);
    int rc = SQ_OK;
    if (b) {
        if (mods[ROLE_QC]) arr->metas_stale = true;
        void *h[6];
        for (int r = 0; r < 6; r++) h[r] = mods[r] ? ((Collector *)mods[r])->h : nullptr;
        Py_BEGIN_ALLOW_THREADS
        rc = sq_fused_add(g_ctx, b, (sq_qc *)h[ROLE_QC], (sq_pertile *)h[ROLE_PT], (sq_overrep *)h[ROLE_OV],
                          (sq_nanostats *)h[ROLE_NS], (sq_adapters *)h[ROLE_AD], (sq_dedup *)h[ROLE_DD]);
        Py_END_ALLOW_THREADS
    }
    for (int r = 0; r < 6; r++) Py_XDECREF(mods[r]);
    Py_DECREF((PyObject *)arr);
    if (!b) return -1;
    if (rc != SQ_OK) {
        raise_sq(rc, "sq_fused_add");
        return -1;
    }
    return 0;
}
int defer_add(int role, PyObject *mod, ArrayView *arr) {
    if (g_pending.array) {
        // NanoStats copies the error sum QCMetrics left in the array (:5314): if it was fed before
        // QCMetrics, keep that order
        const bool clash = g_pending.array != (PyObject *)arr || g_pending.mods[role] ||
                           (role == ROLE_QC && g_pending.mods[ROLE_NS]);
        if (clash && flush_pending() < 0) return -1;
    }
    if (!g_pending.array) {
        Py_INCREF((PyObject *)arr);
        g_pending.array = (PyObject *)arr;
    }
    Py_INCREF(mod);
    g_pending.mods[role] = mod;
    return 0;
}
// (a collector with an add still pending cannot go away: the pending list owns a reference)
template <void (*DESTROY)(void *)>
void Collector_dealloc(Collector *self) {
    PyTypeObject *tp = Py_TYPE(self);
    if (self->h) DESTROY(self->h);
    tp->tp_free((PyObject *)self);
    Py_DECREF(tp);
}

// ---------------------------------------------------------------------------------------------
// QCMetrics (reference :1786-2385)
// ---------------------------------------------------------------------------------------------
void qc_destroy(void *h) { sq_qc_destroy((sq_qc *)h); }
PyObject *QC_new(PyTypeObject *type, PyObject *args, PyObject *kwargs) {
    static const char *kw[] = {"end_anchor_length", nullptr};
    Py_ssize_t ea = 100;
    if (!PyArg_ParseTupleAndKeywords(args, kwargs, "|n:QCMetrics", (char **)kw, &ea)) return nullptr;
    if (ea < 0 || (uint64_t)ea > 0xFFFFFFFFULL) {
        PyErr_Format(PyExc_ValueError, "end_anchor_length must be between 0 and %zd, got %zd", (Py_ssize_t)0xFFFFFFFFLL, ea);
        return nullptr;
    }
    sq_ctx *ctx = ctx_get();
    if (!ctx) return nullptr;
    Collector *self = (Collector *)type->tp_alloc(type, 0);
    if (!self) return nullptr;
    int rc = sq_qc_create(ctx, (uint64_t)ea, (sq_qc **)&self->h);
    if (rc != SQ_OK) {
        self->h = nullptr;
        Py_DECREF(self);
        return raise_sq(rc, "sq_qc_create");
    }
    return (PyObject *)self;
}
int QC_sync(Collector *self, sq_qc_info *info) {
    if (flush_pending() < 0) return -1;
    int rc;
    Py_BEGIN_ALLOW_THREADS
    rc = sq_qc_sync((sq_qc *)self->h, info);
    Py_END_ALLOW_THREADS
    SQ_CHECK_INT(rc, "sq_qc_sync");
    if (info->bad_phred) {
        PyErr_Format(PyExc_ValueError, "Not a valid phred character: %c", (int)info->bad_phred_char);
        return -1;
    }
    return 0;
}
PyObject *QC_add_record_array(Collector *self, PyObject *o) {
    ArrayView *arr = check_array(o, "record_array");
    if (!arr) return nullptr;
    if (arr->n && defer_add(ROLE_QC, (PyObject *)self, arr) < 0) return nullptr;
    Py_RETURN_NONE;
}
PyObject *QC_add_read(Collector *self, PyObject *o) {
    RecordView *v = check_read(o);
    if (!v) return nullptr;
    ArrayView *arr = single_array(v);
    if (!arr) return nullptr;
    sq_qc_info info;
    int rc = defer_add(ROLE_QC, (PyObject *)self, arr);
    if (rc == 0) rc = QC_sync(self, &info);
    // QCMetrics stores the ordered error sum back into the record (:2126)
    if (rc == 0) rc = ArrayView_fetch_metas(arr);
    if (rc == 0) v->err = arr->metas[0].err_sum;
    Py_DECREF((PyObject *)arr);
    if (rc < 0) return nullptr;
    Py_RETURN_NONE;
}
PyObject *QC_table(Collector *self, int which) {
    sq_qc_info info;
    if (QC_sync(self, &info) < 0) return nullptr;
    const size_t sizes[6] = {(size_t)info.max_length * 5, (size_t)info.max_length * 12, (size_t)info.end_anchor_length * 5,
                             (size_t)info.end_anchor_length * 12, 101, 94};
    std::vector<uint64_t> out(sizes[which] ? sizes[which] : 1, 0);
    uint64_t *p[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    p[which] = out.data();
    SQ_CHECK(sq_qc_read((sq_qc *)self->h, p[0], p[1], p[2], p[3], p[4], p[5]), "sq_qc_read");
    return u64_array(out.data(), sizes[which]);
}
PyObject *QC_base(Collector *s, PyObject *) { return QC_table(s, 0); }
PyObject *QC_phred(Collector *s, PyObject *) { return QC_table(s, 1); }
PyObject *QC_ea_base(Collector *s, PyObject *) { return QC_table(s, 2); }
PyObject *QC_ea_phred(Collector *s, PyObject *) { return QC_table(s, 3); }
PyObject *QC_gc(Collector *s, PyObject *) { return QC_table(s, 4); }
PyObject *QC_scores(Collector *s, PyObject *) { return QC_table(s, 5); }
PyObject *QC_get(Collector *self, void *closure) {
    sq_qc_info info;
    if (QC_sync(self, &info) < 0) return nullptr;
    const intptr_t k = (intptr_t)closure;
    return PyLong_FromUnsignedLongLong(k == 0 ? info.max_length : k == 1 ? info.number_of_reads : info.end_anchor_length);
}
// extension of the B200 build: aggregate(data_ranges, count_thresholds=()) -> dict, see sequali_b200/report.py
PyObject *QC_aggregate(Collector *self, PyObject *args) {
    PyObject *ranges_obj = nullptr, *thr_obj = nullptr;
    if (!PyArg_ParseTuple(args, "O|O:aggregate", &ranges_obj, &thr_obj)) return nullptr;
    std::vector<uint64_t> flat, thr;
    if (u64_from_iterable(ranges_obj, true, flat) < 0) return nullptr;
    if (thr_obj && u64_from_iterable(thr_obj, false, thr) < 0) return nullptr;
    sq_qc_info info;
    if (QC_sync(self, &info) < 0) return nullptr;
    const size_t n = flat.size() / 2;
    std::vector<uint64_t> starts(n + 1), stops(n + 1), base(n * 5 + 1), phred(n * 12 + 1), lengths(n + 1);
    for (size_t i = 0; i < n; i++) starts[i] = flat[2 * i], stops[i] = flat[2 * i + 1];
    sq_qc_length_summary sum;
    SQ_CHECK(sq_qc_aggregate((sq_qc *)self->h, starts.data(), stops.data(), n, base.data(), phred.data(), lengths.data(),
                             thr.data(), thr.size(), info.number_of_reads, &sum),
             "sq_qc_aggregate");
    return dict_of({{"base_matrix", u64_array(base.data(), n * 5)},
                    {"phred_matrix", u64_array(phred.data(), n * 12)},
                    {"length_counts", u64_list(lengths.data(), n)},
                    {"total_bases", PyLong_FromUnsignedLongLong(sum.total_bases)},
                    {"minimum_length", PyLong_FromUnsignedLongLong(sum.minimum_length)},
                    {"n50", PyLong_FromUnsignedLongLong(sum.n50)},
                    {"n90", PyLong_FromUnsignedLongLong(sum.n90)},
                    {"threshold_lengths", u64_list(sum.threshold_lengths, thr.size())}});
}
PyMethodDef QC_methods[] = {
    {"add_read", (PyCFunction)QC_add_read, METH_O, "Add a read to the count metrics."},
    {"add_record_array", (PyCFunction)QC_add_record_array, METH_O, "Add a record_array to the count metrics."},
    {"base_count_table", (PyCFunction)QC_base, METH_NOARGS, "array.array('Q') of max_length * NUMBER_OF_NUCS counts"},
    {"phred_count_table", (PyCFunction)QC_phred, METH_NOARGS, "array.array('Q') of max_length * NUMBER_OF_PHREDS counts"},
    {"end_anchored_base_count_table", (PyCFunction)QC_ea_base, METH_NOARGS, "end anchored base counts"},
    {"end_anchored_phred_count_table", (PyCFunction)QC_ea_phred, METH_NOARGS, "end anchored phred counts"},
    {"gc_content", (PyCFunction)QC_gc, METH_NOARGS, "array.array('Q') of 101 GC percentage counts"},
    {"phred_scores", (PyCFunction)QC_scores, METH_NOARGS, "array.array('Q') of PHRED_MAX + 1 mean phred counts"},
    {"aggregate", (PyCFunction)QC_aggregate, METH_VARARGS,
     "aggregate(data_ranges, count_thresholds=()) -> dict: the report's sums over position ranges and its length\n"
     "distribution walk, computed on the device tables (extension of the B200 build)"},
    {nullptr, nullptr, 0, nullptr}};
PyGetSetDef QC_getset[] = {{"max_length", (getter)QC_get, nullptr, "length of the longest read", (void *)0},
                           {"number_of_reads", (getter)QC_get, nullptr, "number of reads processed", (void *)1},
                           {"end_anchor_length", (getter)QC_get, nullptr, "length of the end anchored tables", (void *)2},
                           {nullptr, nullptr, nullptr, nullptr, nullptr}};
PyType_Slot QC_slots[] = {{Py_tp_new, (void *)QC_new},
                          {Py_tp_dealloc, (void *)Collector_dealloc<qc_destroy>},
                          {Py_tp_methods, QC_methods},
                          {Py_tp_getset, QC_getset},
                          {0, nullptr}};
PyType_Spec QC_spec = {"_qc.QCMetrics", sizeof(Collector), 0, Py_TPFLAGS_DEFAULT, QC_slots};

// ---------------------------------------------------------------------------------------------
// AdapterCounter (reference :2406-2969)
// ---------------------------------------------------------------------------------------------
struct Adapters {
    PyObject_HEAD
    void *h;
    PyObject *adapters;  // tuple of str
};
PyObject *AD_new(PyTypeObject *type, PyObject *args, PyObject *kwargs) {
    static const char *kw[] = {"adapters", nullptr};
    PyObject *it = nullptr;
    if (!PyArg_ParseTupleAndKeywords(args, kwargs, "O:AdapterCounter", (char **)kw, &it)) return nullptr;
    PyObject *tup = PySequence_Tuple(it);
    if (!tup) return nullptr;
    const Py_ssize_t n = PyTuple_GET_SIZE(tup);
    if (n < 1) {
        Py_DECREF(tup);
        PyErr_SetString(PyExc_ValueError, "At least one adapter is expected");
        return nullptr;
    }
    std::vector<const char *> ptrs;
    for (Py_ssize_t i = 0; i < n; i++) {
        PyObject *a = PyTuple_GET_ITEM(tup, i);
        if (!PyUnicode_CheckExact(a)) {
            PyErr_Format(PyExc_TypeError, "All adapter sequences must be of type str, got %R, for %R", (PyObject *)Py_TYPE(a), a);
            Py_DECREF(tup);
            return nullptr;
        }
        if (!PyUnicode_IS_COMPACT_ASCII(a)) {
            PyErr_Format(PyExc_ValueError, "Adapter must contain only ASCII characters: %R", a);
            Py_DECREF(tup);
            return nullptr;
        }
        if (PyUnicode_GET_LENGTH(a) > 64) {
            PyErr_Format(PyExc_ValueError, "Maximum adapter size is %d, got %zd for %R", 64, PyUnicode_GET_LENGTH(a), a);
            Py_DECREF(tup);
            return nullptr;
        }
        ptrs.push_back(PyUnicode_AsUTF8(a));
    }
    sq_ctx *ctx = ctx_get();
    Adapters *self = ctx ? (Adapters *)type->tp_alloc(type, 0) : nullptr;
    if (!self) {
        Py_DECREF(tup);
        return nullptr;
    }
    self->adapters = tup;
    int rc = sq_adapters_create(ctx, ptrs.data(), (uint64_t)n, (sq_adapters **)&self->h);
    if (rc != SQ_OK) {
        self->h = nullptr;
        Py_DECREF(self);
        return raise_sq(rc, "sq_adapters_create");
    }
    return (PyObject *)self;
}
void AD_dealloc(Adapters *self) {
    PyTypeObject *tp = Py_TYPE(self);
    if (self->h) sq_adapters_destroy((sq_adapters *)self->h);
    Py_XDECREF(self->adapters);
    tp->tp_free((PyObject *)self);
    Py_DECREF(tp);
}
int AD_sync(Adapters *self, uint64_t *n, uint64_t *ml) {
    if (flush_pending() < 0) return -1;
    int rc;
    Py_BEGIN_ALLOW_THREADS
    rc = sq_adapters_sync((sq_adapters *)self->h, n, ml);
    Py_END_ALLOW_THREADS
    SQ_CHECK_INT(rc, "sq_adapters_sync");
    return 0;
}
PyObject *AD_add_record_array(Adapters *self, PyObject *o) {
    ArrayView *arr = check_array(o, "record_array");
    if (!arr) return nullptr;
    if (arr->n && defer_add(ROLE_AD, (PyObject *)self, arr) < 0) return nullptr;
    Py_RETURN_NONE;
}
PyObject *AD_add_read(Adapters *self, PyObject *o) {
    RecordView *v = check_read(o);
    ArrayView *arr = v ? single_array(v) : nullptr;
    if (!arr) return nullptr;
    uint64_t n, ml;
    int rc = defer_add(ROLE_AD, (PyObject *)self, arr);
    if (rc == 0) rc = AD_sync(self, &n, &ml);
    Py_DECREF((PyObject *)arr);
    if (rc < 0) return nullptr;
    Py_RETURN_NONE;
}
PyObject *AD_get_counts(Adapters *self, PyObject *) {
    uint64_t n, ml;
    if (AD_sync(self, &n, &ml) < 0) return nullptr;
    const Py_ssize_t na = PyTuple_GET_SIZE(self->adapters);
    PyObject *out = PyList_New(na);
    if (!out) return nullptr;
    std::vector<uint64_t> f(ml ? ml : 1), r(ml ? ml : 1);
    for (Py_ssize_t i = 0; i < na; i++) {
        int rc = sq_adapters_read((sq_adapters *)self->h, (uint64_t)i, f.data(), r.data());
        if (rc != SQ_OK) {
            Py_DECREF(out);
            return raise_sq(rc, "sq_adapters_read");
        }
        PyObject *fa = u64_array(f.data(), ml), *ra = fa ? u64_array(r.data(), ml) : nullptr;
        PyObject *t = ra ? PyTuple_Pack(3, PyTuple_GET_ITEM(self->adapters, i), fa, ra) : nullptr;
        Py_XDECREF(fa);
        Py_XDECREF(ra);
        if (!t) {
            Py_DECREF(out);
            return nullptr;
        }
        PyList_SET_ITEM(out, i, t);
    }
    return out;
}
PyObject *AD_get(Adapters *self, void *closure) {
    uint64_t n, ml;
    if (AD_sync(self, &n, &ml) < 0) return nullptr;
    return PyLong_FromUnsignedLongLong(closure ? ml : n);
}
PyMethodDef AD_methods[] = {{"add_read", (PyCFunction)AD_add_read, METH_O, "Add a read to the adapter counter."},
                            {"add_record_array", (PyCFunction)AD_add_record_array, METH_O, "Add a record_array."},
                            {"get_counts", (PyCFunction)AD_get_counts, METH_NOARGS, "[(adapter, forward, reverse)]"},
                            {nullptr, nullptr, 0, nullptr}};
PyGetSetDef AD_getset[] = {{"number_of_sequences", (getter)AD_get, nullptr, "number of reads processed", (void *)0},
                           {"max_length", (getter)AD_get, nullptr, "length of the longest read", (void *)1},
                           {nullptr, nullptr, nullptr, nullptr, nullptr}};
PyMemberDef AD_members[] = {{"adapters", T_OBJECT, offsetof(Adapters, adapters), READONLY, "the adapters searched for"},
                            {nullptr, 0, 0, 0, nullptr}};
PyType_Slot AD_slots[] = {{Py_tp_new, (void *)AD_new},         {Py_tp_dealloc, (void *)AD_dealloc},
                          {Py_tp_methods, AD_methods},         {Py_tp_getset, AD_getset},
                          {Py_tp_members, AD_members},         {0, nullptr}};
PyType_Spec AD_spec = {"_qc.AdapterCounter", sizeof(Adapters), 0, Py_TPFLAGS_DEFAULT, AD_slots};

// ---------------------------------------------------------------------------------------------
// PerTileQuality (reference :2975-3397) and NanoStats (:4882-5450) share the skipped_reason logic
// ---------------------------------------------------------------------------------------------
struct Skippable {
    PyObject_HEAD
    void *h;
    PyObject *reason;  // str once the module switched itself off, else NULL
    uint64_t warned;   // NanoStats: pi warnings already issued
};
PyObject *header_reason(const uint8_t *name, uint64_t len) {
    PyObject *b = PyUnicode_DecodeASCII((const char *)name, (Py_ssize_t)len, "replace");
    if (!b) return nullptr;
    PyObject *r = PyUnicode_FromFormat("Can not parse header: %R", b);
    Py_DECREF(b);
    return r;
}
void pt_destroy(void *h) { sq_pertile_destroy((sq_pertile *)h); }
void ns_destroy(void *h) { sq_nanostats_destroy((sq_nanostats *)h); }
template <void (*DESTROY)(void *)>
void Skippable_dealloc(Skippable *self) {
    PyTypeObject *tp = Py_TYPE(self);
    if (self->h) DESTROY(self->h);
    Py_XDECREF(self->reason);
    tp->tp_free((PyObject *)self);
    Py_DECREF(tp);
}
PyObject *PT_new(PyTypeObject *type, PyObject *args, PyObject *kwargs) {
    static const char *kw[] = {nullptr};
    if (!PyArg_ParseTupleAndKeywords(args, kwargs, ":PerTileQuality", (char **)kw)) return nullptr;
    sq_ctx *ctx = ctx_get();
    Skippable *self = ctx ? (Skippable *)type->tp_alloc(type, 0) : nullptr;
    if (!self) return nullptr;
    self->reason = nullptr;
    int rc = sq_pertile_create(ctx, (sq_pertile **)&self->h);
    if (rc != SQ_OK) {
        self->h = nullptr;
        Py_DECREF(self);
        return raise_sq(rc, "sq_pertile_create");
    }
    return (PyObject *)self;
}
int PT_sync(Skippable *self, sq_pertile_info *info) {
    if (flush_pending() < 0) return -1;
    int rc;
    Py_BEGIN_ALLOW_THREADS
    rc = sq_pertile_sync((sq_pertile *)self->h, info);
    Py_END_ALLOW_THREADS
    SQ_CHECK_INT(rc, "sq_pertile_sync");
    if (info->bad_phred) {
        PyErr_Format(PyExc_ValueError, "Not a valid phred character: %c", (int)info->bad_phred_char);
        return -1;
    }
    if (info->skipped && !self->reason) {
        std::vector<uint8_t> buf(1 << 16);
        uint64_t len = 0;
        SQ_CHECK_INT(sq_pertile_skipped_name((sq_pertile *)self->h, buf.data(), buf.size(), &len), "sq_pertile_skipped_name");
        self->reason = header_reason(buf.data(), len);
        if (!self->reason) return -1;
    }
    return 0;
}
PyObject *PT_add_record_array(Skippable *self, PyObject *o) {
    if (self->reason) Py_RETURN_NONE;
    ArrayView *arr = check_array(o, "record_array");
    if (!arr) return nullptr;
    if (arr->n && defer_add(ROLE_PT, (PyObject *)self, arr) < 0) return nullptr;
    Py_RETURN_NONE;
}
PyObject *PT_add_read(Skippable *self, PyObject *o) {
    if (self->reason) Py_RETURN_NONE;
    RecordView *v = check_read(o);
    ArrayView *arr = v ? single_array(v) : nullptr;
    if (!arr) return nullptr;
    sq_pertile_info info;
    int rc = defer_add(ROLE_PT, (PyObject *)self, arr);
    if (rc == 0) rc = PT_sync(self, &info);
    Py_DECREF((PyObject *)arr);
    if (rc < 0) return nullptr;
    Py_RETURN_NONE;
}
PyObject *PT_get_tile_counts(Skippable *self, PyObject *) {
    sq_pertile_info info;
    if (PT_sync(self, &info) < 0) return nullptr;
    const uint64_t nt = info.n_tiles, ml = info.max_length;
    std::vector<uint64_t> ids(nt ? nt : 1), cnt(nt && ml ? nt * ml : 1);
    std::vector<double> err(nt && ml ? nt * ml : 1);
    if (nt) SQ_CHECK(sq_pertile_read((sq_pertile *)self->h, ids.data(), err.data(), cnt.data()), "sq_pertile_read");
    PyObject *out = PyList_New((Py_ssize_t)nt);
    if (!out) return nullptr;
    for (uint64_t t = 0; t < nt; t++) {
        PyObject *el = PyList_New((Py_ssize_t)ml), *cl = el ? PyList_New((Py_ssize_t)ml) : nullptr;
        bool ok = cl != nullptr;
        // counts[j] = reads of the tile longer than j: one int object per distinct value
        PyObject *last = nullptr;
        uint64_t last_v = 0;
        for (uint64_t j = 0; ok && j < ml; j++) {
            PyObject *f = PyFloat_FromDouble(err[t * ml + j]);
            const uint64_t c = cnt[t * ml + j];
            if (!last || c != last_v) {
                Py_XDECREF(last);
                last = PyLong_FromUnsignedLongLong(c);
                last_v = c;
            }
            if (!f || !last) {
                Py_XDECREF(f);
                ok = false;
                break;
            }
            PyList_SET_ITEM(el, (Py_ssize_t)j, f);
            Py_INCREF(last);
            PyList_SET_ITEM(cl, (Py_ssize_t)j, last);
        }
        Py_XDECREF(last);
        PyObject *tid = ok ? PyLong_FromUnsignedLongLong(ids[t]) : nullptr;
        PyObject *tup = tid ? PyTuple_Pack(3, tid, el, cl) : nullptr;
        Py_XDECREF(tid);
        Py_XDECREF(el);
        Py_XDECREF(cl);
        if (!tup) {
            Py_DECREF(out);
            return nullptr;
        }
        PyList_SET_ITEM(out, (Py_ssize_t)t, tup);
    }
    return out;
}
PyObject *PT_get(Skippable *self, void *closure) {
    sq_pertile_info info;
    if (PT_sync(self, &info) < 0) return nullptr;
    const intptr_t k = (intptr_t)closure;
    if (k == 2) {
        PyObject *r = self->reason ? self->reason : Py_None;
        Py_INCREF(r);
        return r;
    }
    return PyLong_FromUnsignedLongLong(k == 0 ? info.max_length : info.number_of_reads);
}
PyMethodDef PT_methods[] = {{"add_read", (PyCFunction)PT_add_read, METH_O, "Add a read to the PerTileQuality Metrics."},
                            {"add_record_array", (PyCFunction)PT_add_record_array, METH_O, "Add a record_array."},
                            {"get_tile_counts", (PyCFunction)PT_get_tile_counts, METH_NOARGS,
                             "[(tile, summed errors per position, counts per position)]"},
                            {nullptr, nullptr, 0, nullptr}};
PyGetSetDef PT_getset[] = {{"max_length", (getter)PT_get, nullptr, "length of the longest read", (void *)0},
                           {"number_of_reads", (getter)PT_get, nullptr, "number of reads processed", (void *)1},
                           {"skipped_reason", (getter)PT_get, nullptr, "why the module switched itself off, or None", (void *)2},
                           {nullptr, nullptr, nullptr, nullptr, nullptr}};
PyType_Slot PT_slots[] = {{Py_tp_new, (void *)PT_new},
                          {Py_tp_dealloc, (void *)Skippable_dealloc<pt_destroy>},
                          {Py_tp_methods, PT_methods},
                          {Py_tp_getset, PT_getset},
                          {0, nullptr}};
PyType_Spec PT_spec = {"_qc.PerTileQuality", sizeof(Skippable), 0, Py_TPFLAGS_DEFAULT, PT_slots};

// ---------------------------------------------------------------------------------------------
// OverrepresentedSequences (reference :3435-4236)
// ---------------------------------------------------------------------------------------------
struct Overrep {
    PyObject_HEAD
    void *h;
    Py_ssize_t max_unique_fragments, fragment_length, sample_every;
    uint64_t warned;
};
PyObject *OV_new(PyTypeObject *type, PyObject *args, PyObject *kwargs) {
    static const char *kw[] = {"max_unique_fragments", "fragment_length", "sample_every", "bases_from_start",
                               "bases_from_end", nullptr};
    Py_ssize_t max_unique = 5000000, k = 21, every = 8, from_start = 100, from_end = 100;
    if (!PyArg_ParseTupleAndKeywords(args, kwargs, "|nnnnn:OverrepresentedSequences", (char **)kw, &max_unique, &k, &every,
                                     &from_start, &from_end))
        return nullptr;
    if (max_unique < 1) {
        PyErr_Format(PyExc_ValueError, "max_unique_fragments should be at least 1, got: %zd", max_unique);
        return nullptr;
    }
    if ((k & 1) == 0 || k > 31 || k < 3) {
        PyErr_Format(PyExc_ValueError, "fragment_length must be between 3 and 31 and be an uneven number, got: %zd", k);
        return nullptr;
    }
    if (every < 1) {
        PyErr_Format(PyExc_ValueError, "sample_every must be 1 or greater. Got %zd", every);
        return nullptr;
    }
    sq_ctx *ctx = ctx_get();
    Overrep *self = ctx ? (Overrep *)type->tp_alloc(type, 0) : nullptr;
    if (!self) return nullptr;
    self->max_unique_fragments = max_unique;
    self->fragment_length = k;
    self->sample_every = every;
    self->warned = 0;
    int rc = sq_overrep_create(ctx, (uint64_t)max_unique, (uint32_t)k, (uint64_t)every, (int64_t)from_start, (int64_t)from_end,
                               (sq_overrep **)&self->h);
    if (rc != SQ_OK) {
        self->h = nullptr;
        Py_DECREF(self);
        return raise_sq(rc, "sq_overrep_create");
    }
    return (PyObject *)self;
}
void OV_dealloc(Overrep *self) {
    PyTypeObject *tp = Py_TYPE(self);
    if (self->h) sq_overrep_destroy((sq_overrep *)self->h);
    tp->tp_free((PyObject *)self);
    Py_DECREF(tp);
}
int OV_sync(Overrep *self, sq_overrep_info *info, ArrayView *source) {
    if (flush_pending() < 0) return -1;
    int rc;
    Py_BEGIN_ALLOW_THREADS
    rc = sq_overrep_sync((sq_overrep *)self->h, info);
    Py_END_ALLOW_THREADS
    SQ_CHECK_INT(rc, "sq_overrep_sync");
    if (info->warn_records > self->warned) {
        self->warned = info->warn_records;
        PyObject *culprit = nullptr;
        if (source && source->n == 1) {
            PyObject *rec = ArrayView_item(source, 0);
            PyObject *seq = rec ? RecordView_sequence((RecordView *)rec, nullptr) : nullptr;
            Py_XDECREF(rec);
            if (!seq) return -1;
            culprit = PyObject_Repr(seq);
            Py_DECREF(seq);
            if (!culprit) return -1;
        }
        if (!culprit) culprit = PyUnicode_FromString("");
        if (!culprit) return -1;
        int w = PyErr_WarnFormat(PyExc_UserWarning, 1, "Sequence contains a chacter that is not A, C, G, T or N: %U", culprit);
        Py_DECREF(culprit);
        if (w < 0) return -1;
    }
    return 0;
}
PyObject *OV_add_record_array(Overrep *self, PyObject *o) {
    ArrayView *arr = check_array(o, "record_array");
    if (!arr) return nullptr;
    if (arr->n && defer_add(ROLE_OV, (PyObject *)self, arr) < 0) return nullptr;
    Py_RETURN_NONE;
}
PyObject *OV_add_read(Overrep *self, PyObject *o) {
    RecordView *v = check_read(o);
    ArrayView *arr = v ? single_array(v) : nullptr;
    if (!arr) return nullptr;
    sq_overrep_info info;
    int rc = defer_add(ROLE_OV, (PyObject *)self, arr);
    if (rc == 0) rc = OV_sync(self, &info, arr);
    Py_DECREF((PyObject *)arr);
    if (rc < 0) return nullptr;
    Py_RETURN_NONE;
}
PyObject *kmer_string(uint64_t kmer, Py_ssize_t k) {
    char buf[32];
    for (Py_ssize_t i = 0; i < k; i++) buf[i] = "ACGT"[(kmer >> (2 * (k - 1 - i))) & 3];
    return PyUnicode_FromStringAndSize(buf, k);
}
PyObject *OV_sequence_counts(Overrep *self, PyObject *) {
    sq_overrep_info info;
    if (OV_sync(self, &info, nullptr) < 0) return nullptr;
    const uint64_t n = info.collected_unique_fragments;
    std::vector<uint64_t> km(n ? n : 1);
    std::vector<uint32_t> ct(n ? n : 1);
    uint64_t got = 0;
    SQ_CHECK(sq_overrep_read((sq_overrep *)self->h, km.data(), ct.data(), &got), "sq_overrep_read");
    PyObject *d = PyDict_New();
    if (!d) return nullptr;
    for (uint64_t i = 0; i < got; i++) {
        PyObject *key = kmer_string(km[i], self->fragment_length), *val = key ? PyLong_FromUnsignedLong(ct[i]) : nullptr;
        const int rc = val ? PyDict_SetItem(d, key, val) : -1;
        Py_XDECREF(key);
        Py_XDECREF(val);
        if (rc < 0) {
            Py_DECREF(d);
            return nullptr;
        }
    }
    return d;
}
PyObject *OV_overrepresented(Overrep *self, PyObject *args, PyObject *kwargs) {
    static const char *kw[] = {"threshold_fraction", "min_threshold", "max_threshold", nullptr};
    double fraction = 0.0001;
    Py_ssize_t min_t = 1, max_t = PY_SSIZE_T_MAX;
    if (!PyArg_ParseTupleAndKeywords(args, kwargs, "|dnn:overrepresented_sequences", (char **)kw, &fraction, &min_t, &max_t))
        return nullptr;
    if (fraction < 0.0 || fraction > 1.0) {
        PyObject *f = PyFloat_FromDouble(fraction);
        PyErr_Format(PyExc_ValueError, "threshold_fraction must be between 0.0 and 1.0 got, %R", f);
        Py_XDECREF(f);
        return nullptr;
    }
    if (min_t < 1) {
        PyErr_Format(PyExc_ValueError, "min_threshold must be at least 1, got %zd", min_t);
        return nullptr;
    }
    if (max_t < 1) {
        PyErr_Format(PyExc_ValueError, "max_threshold must be at least 1, got %zd", max_t);
        return nullptr;
    }
    sq_overrep_info info;
    if (OV_sync(self, &info, nullptr) < 0) return nullptr;
    const uint64_t sampled = info.sampled_sequences;
    Py_ssize_t hits = (Py_ssize_t)ceil(fraction * (double)sampled);
    hits = std::min(max_t, std::max(min_t, hits));
    // the table is filtered on the device; only the hits cross PCIe
    uint64_t cap = 4096, got = 0;
    std::vector<uint64_t> km;
    std::vector<uint32_t> ct;
    for (;;) {
        km.assign(cap, 0);
        ct.assign(cap, 0);
        SQ_CHECK(sq_overrep_read_min((sq_overrep *)self->h, (uint32_t)std::min<Py_ssize_t>(hits, 0xFFFFFFFFLL), km.data(),
                                     ct.data(), cap, &got),
                 "sq_overrep_read_min");
        if (got <= cap) break;
        cap = got;
    }
    struct Hit { uint32_t count; std::string seq; };
    std::vector<Hit> v;
    v.reserve(got);
    for (uint64_t i = 0; i < got; i++) {
        std::string s((size_t)self->fragment_length, 'A');
        for (Py_ssize_t j = 0; j < self->fragment_length; j++)
            s[j] = "ACGT"[(km[i] >> (2 * (self->fragment_length - 1 - j))) & 3];
        v.push_back({ct[i], s});
    }
    std::sort(v.begin(), v.end(), [](const Hit &a, const Hit &b) { return a.count != b.count ? a.count > b.count : a.seq > b.seq; });
    PyObject *out = PyList_New((Py_ssize_t)v.size());
    if (!out) return nullptr;
    for (size_t i = 0; i < v.size(); i++) {
        PyObject *t = Py_BuildValue("(kds#)", (unsigned long)v[i].count, (double)v[i].count / (double)sampled, v[i].seq.data(),
                                    (Py_ssize_t)v[i].seq.size());
        if (!t) {
            Py_DECREF(out);
            return nullptr;
        }
        PyList_SET_ITEM(out, (Py_ssize_t)i, t);
    }
    return out;
}
PyObject *OV_get(Overrep *self, void *closure) {
    sq_overrep_info info;
    if (OV_sync(self, &info, nullptr) < 0) return nullptr;
    const intptr_t k = (intptr_t)closure;
    return PyLong_FromUnsignedLongLong(k == 0 ? info.number_of_sequences : k == 1 ? info.sampled_sequences
                                       : k == 2 ? info.collected_unique_fragments : info.total_fragments);
}
PyMethodDef OV_methods[] = {
    {"add_read", (PyCFunction)OV_add_read, METH_O, "Add a read to the overrepresented sequences."},
    {"add_record_array", (PyCFunction)OV_add_record_array, METH_O, "Add a record_array."},
    {"sequence_counts", (PyCFunction)OV_sequence_counts, METH_NOARGS, "{fragment: count}"},
    {"overrepresented_sequences", (PyCFunction)OV_overrepresented, METH_VARARGS | METH_KEYWORDS,
     "[(count, fraction, sequence)] above the threshold, most frequent first"},
    {nullptr, nullptr, 0, nullptr}};
PyGetSetDef OV_getset[] = {{"number_of_sequences", (getter)OV_get, nullptr, "reads submitted", (void *)0},
                           {"sampled_sequences", (getter)OV_get, nullptr, "reads sampled", (void *)1},
                           {"collected_unique_fragments", (getter)OV_get, nullptr, "distinct fragments stored", (void *)2},
                           {"total_fragments", (getter)OV_get, nullptr, "fragments looked at", (void *)3},
                           {nullptr, nullptr, nullptr, nullptr, nullptr}};
PyMemberDef OV_members[] = {
    {"max_unique_fragments", T_PYSSIZET, offsetof(Overrep, max_unique_fragments), READONLY, "table capacity"},
    {"fragment_length", T_PYSSIZET, offsetof(Overrep, fragment_length), READONLY, "fragment length"},
    {"sample_every", T_PYSSIZET, offsetof(Overrep, sample_every), READONLY, "one in this many reads is sampled"},
    {nullptr, 0, 0, 0, nullptr}};
PyType_Slot OV_slots[] = {{Py_tp_new, (void *)OV_new},   {Py_tp_dealloc, (void *)OV_dealloc},
                          {Py_tp_methods, OV_methods},   {Py_tp_getset, OV_getset},
                          {Py_tp_members, OV_members},   {0, nullptr}};
PyType_Spec OV_spec = {"_qc.OverrepresentedSequences", sizeof(Overrep), 0, Py_TPFLAGS_DEFAULT, OV_slots};

// ---------------------------------------------------------------------------------------------
// DedupEstimator (reference :4277-4802)
// ---------------------------------------------------------------------------------------------
struct Dedup {
    PyObject_HEAD
    void *h;
    Py_ssize_t front_sequence_length, back_sequence_length, front_sequence_offset, back_sequence_offset;
};
PyObject *DD_new(PyTypeObject *type, PyObject *args, PyObject *kwargs) {
    static const char *kw[] = {"max_stored_fingerprints", "front_sequence_length", "back_sequence_length",
                               "front_sequence_offset", "back_sequence_offset", nullptr};
    Py_ssize_t max_stored = 1000000, fl = 8, bl = 8, fo = 64, bo = 64;
    if (!PyArg_ParseTupleAndKeywords(args, kwargs, "|n$nnnn:DedupEstimator", (char **)kw, &max_stored, &fl, &bl, &fo, &bo))
        return nullptr;
    if (max_stored < 100) {
        PyErr_Format(PyExc_ValueError, "max_stored_fingerprints must be at least 100, not %zd", max_stored);
        return nullptr;
    }
    const Py_ssize_t vals[4] = {fl, bl, fo, bo};
    for (int i = 0; i < 4; i++)
        if (vals[i] < 0) {
            PyErr_Format(PyExc_ValueError, "%s must be at least 0, got %zd.", kw[i + 1], vals[i]);
            return nullptr;
        }
    if (fl + bl == 0) {
        PyErr_SetString(PyExc_ValueError, "The sum of front_sequence_length and back_sequence_length must be at least 0");
        return nullptr;
    }
    sq_ctx *ctx = ctx_get();
    Dedup *self = ctx ? (Dedup *)type->tp_alloc(type, 0) : nullptr;
    if (!self) return nullptr;
    self->front_sequence_length = fl;
    self->back_sequence_length = bl;
    self->front_sequence_offset = fo;
    self->back_sequence_offset = bo;
    int rc = sq_dedup_create(ctx, (uint64_t)max_stored, (uint64_t)fl, (uint64_t)bl, (uint64_t)fo, (uint64_t)bo,
                             (sq_dedup **)&self->h);
    if (rc != SQ_OK) {
        self->h = nullptr;
        Py_DECREF(self);
        return raise_sq(rc, "sq_dedup_create");
    }
    return (PyObject *)self;
}
void DD_dealloc(Dedup *self) {
    PyTypeObject *tp = Py_TYPE(self);
    if (self->h) sq_dedup_destroy((sq_dedup *)self->h);
    tp->tp_free((PyObject *)self);
    Py_DECREF(tp);
}
int DD_sync(Dedup *self, sq_dedup_info *info) {
    if (flush_pending() < 0) return -1;
    int rc;
    Py_BEGIN_ALLOW_THREADS
    rc = sq_dedup_sync((sq_dedup *)self->h, info);
    Py_END_ALLOW_THREADS
    SQ_CHECK_INT(rc, "sq_dedup_sync");
    return 0;
}
PyObject *DD_add_record_array(Dedup *self, PyObject *o) {
    ArrayView *arr = check_array(o, "record_array");
    if (!arr) return nullptr;
    if (arr->n && defer_add(ROLE_DD, (PyObject *)self, arr) < 0) return nullptr;
    Py_RETURN_NONE;
}
PyObject *pair_add(void *h, bool dedup, ArrayView *a1, ArrayView *a2) {
    if (a1->n != a2->n) {
        PyErr_Format(PyExc_ValueError, "record_array1 and record_array2 must be of the same size. Got %zd and %zd respectively.",
                     (Py_ssize_t)a1->n, (Py_ssize_t)a2->n);
        return nullptr;
    }
    if (a1->n) {
        if (flush_pending() < 0) return nullptr;
        sq_batch *h1 = ArrayView_handle(a1), *h2 = h1 ? ArrayView_handle(a2) : nullptr;
        if (!h2) return nullptr;
        int rc;
        Py_BEGIN_ALLOW_THREADS
        rc = dedup ? sq_dedup_add_pair((sq_dedup *)h, h1, h2) : sq_insert_add_pair((sq_insert *)h, h1, h2);
        Py_END_ALLOW_THREADS
        SQ_CHECK(rc, dedup ? "sq_dedup_add_pair" : "sq_insert_add_pair");
    }
    Py_RETURN_NONE;
}
PyObject *DD_add_record_array_pair(Dedup *self, PyObject *args) {
    PyObject *o1, *o2;
    if (!PyArg_ParseTuple(args, "OO:add_record_array_pair", &o1, &o2)) return nullptr;
    ArrayView *a1 = check_array(o1, "record_array1"), *a2 = a1 ? check_array(o2, "record_array2") : nullptr;
    if (!a2) return nullptr;
    return pair_add(self->h, true, a1, a2);
}
PyObject *DD_add_sequence(Dedup *self, PyObject *s) {
    ArrayView *arr = sequence_array(s, "sequence");
    if (!arr) return nullptr;
    const int rc = defer_add(ROLE_DD, (PyObject *)self, arr);
    Py_DECREF((PyObject *)arr);
    if (rc < 0) return nullptr;
    Py_RETURN_NONE;
}
PyObject *DD_add_sequence_pair(Dedup *self, PyObject *args) {
    PyObject *s1, *s2;
    if (!PyArg_ParseTuple(args, "OO:add_sequence_pair", &s1, &s2)) return nullptr;
    ArrayView *a1 = sequence_array(s1, "sequence"), *a2 = a1 ? sequence_array(s2, "sequence") : nullptr;
    PyObject *r = a2 ? pair_add(self->h, true, a1, a2) : nullptr;
    Py_XDECREF((PyObject *)a1);
    Py_XDECREF((PyObject *)a2);
    return r;
}
PyObject *DD_duplication_counts(Dedup *self, PyObject *) {
    sq_dedup_info info;
    if (DD_sync(self, &info) < 0) return nullptr;
    std::vector<uint64_t> out(info.tracked_sequences ? info.tracked_sequences : 1);
    uint64_t n = 0;
    int rc;
    Py_BEGIN_ALLOW_THREADS
    rc = sq_dedup_read((sq_dedup *)self->h, out.data(), &n);
    Py_END_ALLOW_THREADS
    SQ_CHECK(rc, "sq_dedup_read");
    return u64_array(out.data(), (size_t)std::min<uint64_t>(n, info.tracked_sequences));
}
PyObject *DD_get(Dedup *self, void *closure) {
    sq_dedup_info info;
    if (DD_sync(self, &info) < 0) return nullptr;
    const intptr_t k = (intptr_t)closure;
    return PyLong_FromUnsignedLongLong(k == 0 ? info.modulo_bits : k == 1 ? info.hash_table_size : info.tracked_sequences);
}
PyMethodDef DD_methods[] = {
    {"add_record_array", (PyCFunction)DD_add_record_array, METH_O, "Add a record_array to the dedup estimator."},
    {"add_record_array_pair", (PyCFunction)DD_add_record_array_pair, METH_VARARGS, "Add a pair of record arrays."},
    {"add_sequence", (PyCFunction)DD_add_sequence, METH_O, "Add a sequence to the dedup estimator."},
    {"add_sequence_pair", (PyCFunction)DD_add_sequence_pair, METH_VARARGS, "Add a pair of sequences."},
    {"duplication_counts", (PyCFunction)DD_duplication_counts, METH_NOARGS, "array.array('Q') of the stored counts"},
    {nullptr, nullptr, 0, nullptr}};
PyGetSetDef DD_getset[] = {{"_modulo_bits", (getter)DD_get, nullptr, "sampling level", (void *)0},
                           {"_hash_table_size", (getter)DD_get, nullptr, "slots of the table", (void *)1},
                           {"tracked_sequences", (getter)DD_get, nullptr, "stored fingerprints", (void *)2},
                           {nullptr, nullptr, nullptr, nullptr, nullptr}};
PyMemberDef DD_members[] = {
    {"front_sequence_length", T_PYSSIZET, offsetof(Dedup, front_sequence_length), READONLY, ""},
    {"back_sequence_length", T_PYSSIZET, offsetof(Dedup, back_sequence_length), READONLY, ""},
    {"front_sequence_offset", T_PYSSIZET, offsetof(Dedup, front_sequence_offset), READONLY, ""},
    {"back_sequence_offset", T_PYSSIZET, offsetof(Dedup, back_sequence_offset), READONLY, ""},
    {nullptr, 0, 0, 0, nullptr}};
PyType_Slot DD_slots[] = {{Py_tp_new, (void *)DD_new},   {Py_tp_dealloc, (void *)DD_dealloc},
                          {Py_tp_methods, DD_methods},   {Py_tp_getset, DD_getset},
                          {Py_tp_members, DD_members},   {0, nullptr}};
PyType_Spec DD_spec = {"_qc.DedupEstimator", sizeof(Dedup), 0, Py_TPFLAGS_DEFAULT, DD_slots};

// ---------------------------------------------------------------------------------------------
// NanoStats, NanoporeReadInfo, NanoStatsIterator (reference :4804-5450)
// ---------------------------------------------------------------------------------------------
struct ReadInfo {
    PyObject_HEAD
    sq_nanoinfo info;
};
PyTypeObject *ReadInfoType = nullptr, *NanoIterType = nullptr;
PyObject *ReadInfo_get(ReadInfo *self, void *closure) {
    switch ((intptr_t)closure) {
    case 0: return PyLong_FromLongLong(self->info.start_time);
    case 1: return PyLong_FromLong(self->info.channel_id);
    case 2: return PyLong_FromUnsignedLong(self->info.length);
    case 3: return PyFloat_FromDouble(self->info.cumulative_error_rate);
    case 4: return PyFloat_FromDouble((double)self->info.duration);
    default: return PyLong_FromUnsignedLongLong(self->info.parent_id_hash);
    }
}
PyGetSetDef ReadInfo_getset[] = {{"start_time", (getter)ReadInfo_get, nullptr, "unix UTC timestamp", (void *)0},
                                 {"channel_id", (getter)ReadInfo_get, nullptr, "channel", (void *)1},
                                 {"length", (getter)ReadInfo_get, nullptr, "read length", (void *)2},
                                 {"cumulative_error_rate", (getter)ReadInfo_get, nullptr, "sum of the error rates", (void *)3},
                                 {"duration", (getter)ReadInfo_get, nullptr, "seconds", (void *)4},
                                 {"parent_id_hash", (getter)ReadInfo_get, nullptr, "hash of the parent read id", (void *)5},
                                 {nullptr, nullptr, nullptr, nullptr, nullptr}};
void plain_dealloc(PyObject *self) {
    PyTypeObject *tp = Py_TYPE(self);
    tp->tp_free(self);
    Py_DECREF(tp);
}
PyType_Slot ReadInfo_slots[] = {{Py_tp_dealloc, (void *)plain_dealloc}, {Py_tp_getset, ReadInfo_getset}, {0, nullptr}};
PyType_Spec ReadInfo_spec = {"_qc.NanoporeReadInfo", sizeof(ReadInfo), 0, Py_TPFLAGS_DEFAULT | Py_TPFLAGS_DISALLOW_INSTANTIATION,
                             ReadInfo_slots};
struct NanoIter {
    PyObject_HEAD
    sq_nanoinfo *infos;
    uint64_t n, pos;
};
void NanoIter_dealloc(NanoIter *self) {
    PyTypeObject *tp = Py_TYPE(self);
    free(self->infos);
    tp->tp_free((PyObject *)self);
    Py_DECREF(tp);
}
PyObject *NanoIter_iter(PyObject *self) {
    Py_INCREF(self);
    return self;
}
PyObject *NanoIter_next(NanoIter *self) {
    if (self->pos == self->n) return nullptr;  // StopIteration
    ReadInfo *r = (ReadInfo *)ReadInfoType->tp_alloc(ReadInfoType, 0);
    if (!r) return nullptr;
    r->info = self->infos[self->pos++];
    return (PyObject *)r;
}
PyType_Slot NanoIter_slots[] = {{Py_tp_dealloc, (void *)NanoIter_dealloc},
                                {Py_tp_iter, (void *)NanoIter_iter},
                                {Py_tp_iternext, (void *)NanoIter_next},
                                {0, nullptr}};
PyType_Spec NanoIter_spec = {"_qc.NanoStatsIterator", sizeof(NanoIter), 0, Py_TPFLAGS_DEFAULT | Py_TPFLAGS_DISALLOW_INSTANTIATION,
                             NanoIter_slots};

PyObject *NS_new(PyTypeObject *type, PyObject *args, PyObject *kwargs) {
    static const char *kw[] = {nullptr};
    if (!PyArg_ParseTupleAndKeywords(args, kwargs, ":NanoStats", (char **)kw)) return nullptr;
    sq_ctx *ctx = ctx_get();
    Skippable *self = ctx ? (Skippable *)type->tp_alloc(type, 0) : nullptr;
    if (!self) return nullptr;
    self->reason = nullptr;
    self->warned = 0;
    int rc = sq_nanostats_create(ctx, (sq_nanostats **)&self->h);
    if (rc != SQ_OK) {
        self->h = nullptr;
        Py_DECREF(self);
        return raise_sq(rc, "sq_nanostats_create");
    }
    return (PyObject *)self;
}
int NS_sync(Skippable *self, sq_nanostats_info *info) {
    if (flush_pending() < 0) return -1;
    int rc;
    Py_BEGIN_ALLOW_THREADS
    rc = sq_nanostats_sync((sq_nanostats *)self->h, info);
    Py_END_ALLOW_THREADS
    SQ_CHECK_INT(rc, "sq_nanostats_sync");
    if (info->tag_error) {  // the reference's exceptions for malformed aux data (:5078-5259)
        const int got = (int)(info->tag_error_detail & 0xff);
        static const char *tags[] = {"st", "du", "pi"};
        static const char expected[] = {'Z', 'f', 'Z'};
        const unsigned which = std::min<unsigned>(info->tag_error_detail >> 8, 2);
        switch (info->tag_error) {
        case 2: PyErr_Format(PyExc_ValueError, "Invalid type for array %c", got); break;
        case 3: PyErr_Format(PyExc_ValueError, "Unknown tag type %c", got); break;
        case 4: PyErr_Format(PyExc_RuntimeError, "Wrong tag type for '%s' expected '%c' got '%c'", tags[which], expected[which], got); break;
        case 5:
            PyErr_SetString(PyExc_SystemError, "error return without exception set");
            break;
        default: PyErr_SetString(PyExc_ValueError, "truncated tags");
        }
        return -1;
    }
    if (info->pi_warnings > self->warned) {
        self->warned = info->pi_warnings;
        if (PyErr_WarnFormat(PyExc_UserWarning, 1, "pi tag should have a valid uuid4 format with 36 characters. Counted %u. Skipping tag.",
                             info->pi_first_length) < 0)
            return -1;
    }
    if (info->skipped && !self->reason) {
        std::vector<uint8_t> buf(1 << 16);
        uint64_t len = 0;
        SQ_CHECK_INT(sq_nanostats_skipped_name((sq_nanostats *)self->h, buf.data(), buf.size(), &len),
                     "sq_nanostats_skipped_name");
        self->reason = header_reason(buf.data(), len);
        if (!self->reason) return -1;
    }
    return 0;
}
PyObject *NS_add_record_array(Skippable *self, PyObject *o) {
    ArrayView *arr = check_array(o, "record_array");
    if (!arr) return nullptr;
    if (self->reason) Py_RETURN_NONE;
    if (arr->n && defer_add(ROLE_NS, (PyObject *)self, arr) < 0) return nullptr;
    Py_RETURN_NONE;
}
PyObject *NS_add_read(Skippable *self, PyObject *o) {
    RecordView *v = check_read(o);
    ArrayView *arr = v ? single_array(v) : nullptr;
    if (!arr) return nullptr;
    sq_nanostats_info info;
    int rc = self->reason ? 0 : defer_add(ROLE_NS, (PyObject *)self, arr);
    if (rc == 0) rc = NS_sync(self, &info);
    Py_DECREF((PyObject *)arr);
    if (rc < 0) return nullptr;
    Py_RETURN_NONE;
}
PyObject *NS_iterator(Skippable *self, PyObject *) {
    sq_nanostats_info info;
    if (NS_sync(self, &info) < 0) return nullptr;
    NanoIter *it = (NanoIter *)NanoIterType->tp_alloc(NanoIterType, 0);
    if (!it) return nullptr;
    it->n = info.number_of_reads;
    it->pos = 0;
    it->infos = (sq_nanoinfo *)calloc(it->n ? it->n : 1, sizeof(sq_nanoinfo));
    if (!it->infos) {
        Py_DECREF(it);
        return PyErr_NoMemory();
    }
    if (it->n) {
        int rc = sq_nanostats_read((sq_nanostats *)self->h, it->infos);
        if (rc != SQ_OK) {
            Py_DECREF(it);
            return raise_sq(rc, "sq_nanostats_read");
        }
    }
    return (PyObject *)it;
}
PyObject *NS_get(Skippable *self, void *closure) {
    sq_nanostats_info info;
    if (NS_sync(self, &info) < 0) return nullptr;
    switch ((intptr_t)closure) {
    case 0: return PyLong_FromUnsignedLongLong(info.number_of_reads);
    case 1: return PyLong_FromLongLong(info.minimum_time);
    case 2: return PyLong_FromLongLong(info.maximum_time);
    default: {
        PyObject *r = self->reason ? self->reason : Py_None;
        Py_INCREF(r);
        return r;
    }
    }
}
// extension of the B200 build: report_tables(run_start_time, time_interval, time_slots) -> dict
PyObject *NS_report_tables(Skippable *self, PyObject *args) {
    long long run_start, interval;
    unsigned long long n_slots;
    if (!PyArg_ParseTuple(args, "LLK:report_tables", &run_start, &interval, &n_slots)) return nullptr;
    sq_nanostats_info info;
    if (NS_sync(self, &info) < 0) return nullptr;
    if (n_slots < 1 || n_slots > (1u << 24)) {
        PyErr_SetString(PyExc_ValueError, "time_slots out of range");
        return nullptr;
    }
    std::vector<uint64_t> t_bases(n_slots), t_reads(n_slots), t_active(n_slots), t_quals(n_slots * 12), speeds(81);
    uint64_t parents = 0, n_ch = 0;
    sq_nano_report_error err;
    sq_nanostats *h = (sq_nanostats *)self->h;
    SQ_CHECK(sq_nanostats_report(h, run_start, interval, n_slots, t_bases.data(), t_reads.data(), t_active.data(),
                                 t_quals.data(), speeds.data(), &parents, &n_ch, &err),
             "sq_nanostats_report");
    std::vector<int32_t> ch(n_ch + 1);
    std::vector<uint64_t> ch_bases(n_ch + 1);
    std::vector<double> ch_err(n_ch + 1);
    SQ_CHECK(sq_nanostats_report_channels(h, ch.data(), ch_bases.data(), ch_err.data(), n_ch), "sq_nanostats_report_channels");
    if (err.kind == 1) {
        PyErr_SetString(PyExc_OverflowError, "cannot convert float infinity to integer");
        return nullptr;
    }
    if (err.kind == 2) {
        PyErr_SetString(PyExc_ValueError, "cannot convert float NaN to integer");
        return nullptr;
    }
    if (err.kind == 3) {
        PyErr_SetString(PyExc_IndexError, "list index out of range");
        return nullptr;
    }
    PyObject *quals = PyList_New((Py_ssize_t)n_slots);
    for (size_t i = 0; quals && i < n_slots; i++) {
        PyObject *row = u64_list(t_quals.data() + i * 12, 12);
        if (!row) {
            Py_CLEAR(quals);
            break;
        }
        PyList_SET_ITEM(quals, (Py_ssize_t)i, row);
    }
    PyObject *channels = PyList_New((Py_ssize_t)n_ch), *errors = PyList_New((Py_ssize_t)n_ch);
    for (size_t i = 0; channels && errors && i < n_ch; i++) {
        PyObject *c = PyLong_FromLong(ch[i]), *e = PyFloat_FromDouble(ch_err[i]);
        if (!c || !e) {
            Py_XDECREF(c);
            Py_XDECREF(e);
            Py_CLEAR(channels);
            break;
        }
        PyList_SET_ITEM(channels, (Py_ssize_t)i, c);
        PyList_SET_ITEM(errors, (Py_ssize_t)i, e);
    }
    return dict_of({{"time_bases", u64_list(t_bases.data(), n_slots)},
                    {"time_reads", u64_list(t_reads.data(), n_slots)},
                    {"time_active_channels", u64_list(t_active.data(), n_slots)},
                    {"time_qualities", quals},
                    {"translocation_speed", u64_list(speeds.data(), 81)},
                    {"reads_with_parent", PyLong_FromUnsignedLongLong(parents)},
                    {"channels", channels},
                    {"channel_bases", u64_list(ch_bases.data(), n_ch)},
                    {"channel_cumulative_error", errors}});
}
PyMethodDef NS_methods[] = {{"add_read", (PyCFunction)NS_add_read, METH_O, "Add a read to the NanoStats module."},
                            {"add_record_array", (PyCFunction)NS_add_record_array, METH_O, "Add a record_array."},
                            {"nano_info_iterator", (PyCFunction)NS_iterator, METH_NOARGS, "iterator over NanoporeReadInfo"},
                            {"report_tables", (PyCFunction)NS_report_tables, METH_VARARGS,
                             "report_tables(run_start_time, time_interval, time_slots) -> dict: the per-read loop of the\n"
                             "report's NanoStats module on the device (extension of the B200 build)"},
                            {nullptr, nullptr, 0, nullptr}};
PyGetSetDef NS_getset[] = {{"number_of_reads", (getter)NS_get, nullptr, "reads processed", (void *)0},
                           {"minimum_time", (getter)NS_get, nullptr, "earliest start time", (void *)1},
                           {"maximum_time", (getter)NS_get, nullptr, "latest start time", (void *)2},
                           {"skipped_reason", (getter)NS_get, nullptr, "why the module switched itself off, or None", (void *)3},
                           {nullptr, nullptr, nullptr, nullptr, nullptr}};
PyType_Slot NS_slots[] = {{Py_tp_new, (void *)NS_new},
                          {Py_tp_dealloc, (void *)Skippable_dealloc<ns_destroy>},
                          {Py_tp_methods, NS_methods},
                          {Py_tp_getset, NS_getset},
                          {0, nullptr}};
PyType_Spec NS_spec = {"_qc.NanoStats", sizeof(Skippable), 0, Py_TPFLAGS_DEFAULT, NS_slots};

// ---------------------------------------------------------------------------------------------
// InsertSizeMetrics (reference :5466-5982)
// ---------------------------------------------------------------------------------------------
void is_destroy(void *h) { sq_insert_destroy((sq_insert *)h); }
PyObject *IS_new(PyTypeObject *type, PyObject *args, PyObject *kwargs) {
    static const char *kw[] = {"max_adapters", nullptr};
    Py_ssize_t max_adapters = 10000;
    if (!PyArg_ParseTupleAndKeywords(args, kwargs, "|n:InsertSizeMetrics", (char **)kw, &max_adapters)) return nullptr;
    if (max_adapters < 1) {
        PyErr_Format(PyExc_ValueError, "max_adapters must be at least 1, got %zd", max_adapters);
        return nullptr;
    }
    sq_ctx *ctx = ctx_get();
    Collector *self = ctx ? (Collector *)type->tp_alloc(type, 0) : nullptr;
    if (!self) return nullptr;
    int rc = sq_insert_create(ctx, (uint64_t)max_adapters, (sq_insert **)&self->h);
    if (rc != SQ_OK) {
        self->h = nullptr;
        Py_DECREF(self);
        return raise_sq(rc, "sq_insert_create");
    }
    return (PyObject *)self;
}
int IS_sync(Collector *self, sq_insert_info *info) {
    if (flush_pending() < 0) return -1;
    int rc;
    Py_BEGIN_ALLOW_THREADS
    rc = sq_insert_sync((sq_insert *)self->h, info);
    Py_END_ALLOW_THREADS
    SQ_CHECK_INT(rc, "sq_insert_sync");
    return 0;
}
PyObject *IS_add_record_array_pair(Collector *self, PyObject *args) {
    PyObject *o1, *o2;
    if (!PyArg_ParseTuple(args, "OO:add_record_array_pair", &o1, &o2)) return nullptr;
    ArrayView *a1 = check_array(o1, "record_array1"), *a2 = a1 ? check_array(o2, "record_array2") : nullptr;
    if (!a2) return nullptr;
    return pair_add(self->h, false, a1, a2);
}
PyObject *IS_add_sequence_pair(Collector *self, PyObject *args) {
    PyObject *s1, *s2;
    if (!PyArg_ParseTuple(args, "OO:add_sequence_pair", &s1, &s2)) return nullptr;
    PyObject *both[2] = {s1, s2};
    for (int i = 0; i < 2; i++)
        if (!PyUnicode_Check(both[i])) {
            PyErr_Format(PyExc_TypeError, "InsertSizeMetrics.add_sequence_pair() argument %d must be str, not %s", i + 1,
                         Py_TYPE(both[i])->tp_name);
            return nullptr;
        }
    ArrayView *a1 = sequence_array(s1, "sequence1"), *a2 = a1 ? sequence_array(s2, "sequence2") : nullptr;
    PyObject *r = a2 ? pair_add(self->h, false, a1, a2) : nullptr;
    Py_XDECREF((PyObject *)a1);
    Py_XDECREF((PyObject *)a2);
    return r;
}
PyObject *IS_insert_sizes(Collector *self, PyObject *) {
    sq_insert_info info;
    if (IS_sync(self, &info) < 0) return nullptr;
    std::vector<uint64_t> out(info.max_insert_size + 1);
    SQ_CHECK(sq_insert_read_sizes((sq_insert *)self->h, out.data()), "sq_insert_read_sizes");
    return u64_array(out.data(), out.size());
}
PyObject *IS_adapters(Collector *self, int which) {
    sq_insert_info info;
    if (IS_sync(self, &info) < 0) return nullptr;
    const uint64_t n = which == 0 ? info.entries_read1 : info.entries_read2;
    std::vector<uint8_t> seqs((n ? n : 1) * 32);
    std::vector<uint64_t> cnt(n ? n : 1);
    uint64_t got = 0;
    SQ_CHECK(sq_insert_read_adapters((sq_insert *)self->h, which, seqs.data(), cnt.data(), &got), "sq_insert_read_adapters");
    PyObject *out = PyList_New((Py_ssize_t)got);
    if (!out) return nullptr;
    for (uint64_t i = 0; i < got; i++) {
        PyObject *t = Py_BuildValue("(s#K)", (const char *)&seqs[i * 32 + 1], (Py_ssize_t)seqs[i * 32],
                                    (unsigned long long)cnt[i]);
        if (!t) {
            Py_DECREF(out);
            return nullptr;
        }
        PyList_SET_ITEM(out, (Py_ssize_t)i, t);
    }
    return out;
}
PyObject *IS_adapters1(Collector *s, PyObject *) { return IS_adapters(s, 0); }
PyObject *IS_adapters2(Collector *s, PyObject *) { return IS_adapters(s, 1); }
PyObject *IS_get(Collector *self, void *closure) {
    sq_insert_info info;
    if (IS_sync(self, &info) < 0) return nullptr;
    const intptr_t k = (intptr_t)closure;
    return PyLong_FromUnsignedLongLong(k == 0 ? info.total_reads : k == 1 ? info.number_of_adapters_read1
                                                                          : info.number_of_adapters_read2);
}
PyMethodDef IS_methods[] = {
    {"add_record_array_pair", (PyCFunction)IS_add_record_array_pair, METH_VARARGS, "Add a pair of record arrays."},
    {"add_sequence_pair", (PyCFunction)IS_add_sequence_pair, METH_VARARGS, "Add a pair of sequences."},
    {"insert_sizes", (PyCFunction)IS_insert_sizes, METH_NOARGS, "array.array('Q') of insert size counts"},
    {"adapters_read1", (PyCFunction)IS_adapters1, METH_NOARGS, "[(adapter, count)] of read 1"},
    {"adapters_read2", (PyCFunction)IS_adapters2, METH_NOARGS, "[(adapter, count)] of read 2"},
    {nullptr, nullptr, 0, nullptr}};
PyGetSetDef IS_getset[] = {{"total_reads", (getter)IS_get, nullptr, "pairs processed", (void *)0},
                           {"number_of_adapters_read1", (getter)IS_get, nullptr, "adapters found in read 1", (void *)1},
                           {"number_of_adapters_read2", (getter)IS_get, nullptr, "adapters found in read 2", (void *)2},
                           {nullptr, nullptr, nullptr, nullptr, nullptr}};
PyType_Slot IS_slots[] = {{Py_tp_new, (void *)IS_new},
                          {Py_tp_dealloc, (void *)Collector_dealloc<is_destroy>},
                          {Py_tp_methods, IS_methods},
                          {Py_tp_getset, IS_getset},
                          {0, nullptr}};
PyType_Spec IS_spec = {"_qc.InsertSizeMetrics", sizeof(Collector), 0, Py_TPFLAGS_DEFAULT, IS_slots};

// ---------------------------------------------------------------------------------------------
// FastqParser (reference :889-1244): the host side only stages bytes, the record boundaries are
// found on the device (sq_batch_from_fastq); buffer growth and leftover rules are those of
// FastqParser_create_record_array (:965-1184)
// ---------------------------------------------------------------------------------------------
constexpr Py_ssize_t DEFAULT_FASTQ_BUFFERSIZE = 32 << 20;
constexpr Py_ssize_t DEFAULT_BAM_BUFFERSIZE = 24 << 20;

// Read-ahead of the parsers' __next__: while the caller feeds the collectors with record array k, a helper
// thread reads, copies and scans array k + 1 (readinto needs the GIL; the host->device copy and the
// boundary scan run without it, on the context's parser stream).  This is the double buffering the
// reference gets from xopen's reader threads (src/sequali/util.py:108-123).  Only for read steps of
// 1 MiB and more: the small-buffer call patterns of the tests stay strictly synchronous.
constexpr Py_ssize_t READ_AHEAD_MIN_STEP = 1 << 20;
struct ReadAhead {
    std::mutex mu;
    std::condition_variable cv;
    bool active = false, done = false;
    PyObject *result = nullptr;                                  // new reference, or NULL with the exception below
    PyObject *exc_type = nullptr, *exc_value = nullptr, *exc_tb = nullptr;
};

struct Parser {
    PyObject_HEAD
    PyObject *file;
    PyObject *header;  // BamParser only
    Py_ssize_t read_in_size;
    std::vector<uint8_t> *leftover;
    Pinned bam_buf;    // BamParser: staging buffer kept between calls, leftover at its front
    size_t bam_filled;
    int32_t bam_n_ref;  // reference count of the BAM header
    int direct_fd;      // >= 0: `file` is a plain io.BufferedReader / io.FileIO over a regular file (parser_read)
    struct FileReader *reader;  // FastqParser iterating over such a file: blocks read ahead of the parser
    bool reader_done;           // ... has handed out its last block (or was stopped): the classic path continues
    ReadAhead *ra;
};
int direct_fd_of(PyObject *file);
Py_ssize_t parser_read(Parser *self, uint8_t *dst, Py_ssize_t len);
long long parallel_pread(int fd, uint8_t *dst, size_t len, long long pos, int *err);
void reader_stop(Parser *self, bool seek);

// start `produce(self)` on a helper thread; the thread owns a reference to the parser until it is done
void read_ahead_start(Parser *self, PyObject *(*produce)(Parser *)) {
    ReadAhead *ra = self->ra;
    ra->active = true;
    ra->done = false;
    Py_INCREF((PyObject *)self);
    std::thread([self, ra, produce]() {
        PyGILState_STATE g = PyGILState_Ensure();
        PyObject *r = produce(self);
        PyObject *t = nullptr, *v = nullptr, *tb = nullptr;
        if (!r) PyErr_Fetch(&t, &v, &tb);
        {
            std::lock_guard<std::mutex> lk(ra->mu);
            ra->result = r;
            ra->exc_type = t;
            ra->exc_value = v;
            ra->exc_tb = tb;
            ra->done = true;
        }
        ra->cv.notify_all();
        Py_DECREF((PyObject *)self);  // (may run the parser's dealloc: nothing of it is touched afterwards)
        PyGILState_Release(g);
    }).detach();
}
// wait for the helper thread (GIL released meanwhile) and take what it produced: a new reference, or NULL
// with the exception set (NULL without exception = the producer's plain NULL, e.g. StopIteration)
PyObject *read_ahead_take(Parser *self) {
    ReadAhead *ra = self->ra;
    Py_BEGIN_ALLOW_THREADS
    {
        std::unique_lock<std::mutex> lk(ra->mu);
        ra->cv.wait(lk, [ra] { return ra->done; });
    }
    Py_END_ALLOW_THREADS
    ra->active = false;
    PyObject *r = ra->result;
    ra->result = nullptr;
    if (!r && ra->exc_type) PyErr_Restore(ra->exc_type, ra->exc_value, ra->exc_tb);
    ra->exc_type = ra->exc_value = ra->exc_tb = nullptr;
    return r;
}
void Parser_dealloc(Parser *self) {
    PyTypeObject *tp = Py_TYPE(self);
    // (a helper thread holds a reference while it runs, so none is running here; what it left is dropped)
    if (self->ra) {
        Py_XDECREF(self->ra->result);
        Py_XDECREF(self->ra->exc_type);
        Py_XDECREF(self->ra->exc_value);
        Py_XDECREF(self->ra->exc_tb);
        delete self->ra;
    }
    reader_stop(self, true);
    Py_XDECREF(self->file);
    Py_XDECREF(self->header);
    delete self->leftover;
    self->bam_buf.release();
    tp->tp_free((PyObject *)self);
    Py_DECREF(tp);
}
PyObject *Parser_iter(PyObject *self) {
    Py_INCREF(self);
    return self;
}
PyObject *FQ_new(PyTypeObject *type, PyObject *args, PyObject *kwargs) {
    static const char *kw[] = {"fileobj", "initial_buffersize", nullptr};
    PyObject *file = nullptr, *size_o = Py_None;
    if (!PyArg_ParseTupleAndKeywords(args, kwargs, "O|O:FastqParser", (char **)kw, &file, &size_o)) return nullptr;
    Py_ssize_t size = DEFAULT_FASTQ_BUFFERSIZE;
    if (size_o != Py_None) {
        if (!PyLong_Check(size_o)) {
            PyErr_SetString(PyExc_TypeError, "initial_buffersize must be an integer");
            return nullptr;
        }
        size = PyLong_AsSsize_t(size_o);
        if (size == -1 && PyErr_Occurred()) return nullptr;
    }
    if (size < 1) {
        PyErr_Format(PyExc_ValueError, "initial_buffersize must be at least 1, got %zd", size);
        return nullptr;
    }
    if (!ctx_get()) return nullptr;
    Parser *self = (Parser *)type->tp_alloc(type, 0);
    if (!self) return nullptr;
    Py_INCREF(file);
    self->file = file;
    self->header = nullptr;
    self->read_in_size = size;
    self->leftover = new std::vector<uint8_t>();
    new (&self->bam_buf) Pinned();
    self->bam_filled = 0;
    self->direct_fd = size >= (Py_ssize_t)READ_AHEAD_MIN_STEP ? direct_fd_of(file) : -1;
    self->reader = nullptr;
    self->reader_done = false;
    self->ra = new ReadAhead();
    return (PyObject *)self;
}
// fileobj.readinto(memoryview of [dst, dst + len)) -> bytes read, -1 on error
Py_ssize_t call_readinto(PyObject *file, uint8_t *dst, Py_ssize_t len) {
    PyObject *mv = PyMemoryView_FromMemory((char *)dst, len, PyBUF_WRITE);
    if (!mv) return -1;
    PyObject *res = PyObject_CallMethod(file, "readinto", "O", mv);
    Py_DECREF(mv);
    if (!res) return -1;
    Py_ssize_t got = 0;
    if (res != Py_None) {
        got = PyLong_AsSsize_t(res);
        if (got == -1 && PyErr_Occurred()) {
            Py_DECREF(res);
            return -1;
        }
    }
    Py_DECREF(res);
    if (got < 0 || got > len) {
        PyErr_Format(PyExc_ValueError, "readinto returned %zd for a buffer of %zd bytes", got, len);
        return -1;
    }
    return got;
}
// ---- regular files: the bytes come straight from the descriptor, several threads at a time ----
// `fileobj.readinto` is ONE host thread copying every byte (page cache -> buffer): ~10 GB/s, five times below
// what PCIe takes.  When the object is exactly an io.BufferedReader over an io.FileIO, or an io.FileIO, of a
// regular file (what open(path, "rb") and the reference's xopen give for an uncompressed file), the same bytes
// are fetched with pread() at the object's logical position by up to eight threads and the object is told to
// seek past them -- for the caller nothing differs from a readinto of the same size.  Anything else (gzip
// objects, BytesIO, sockets, subclasses) keeps going through readinto.
constexpr size_t DIRECT_MIN_BYTES = (size_t)4 << 20;  // per thread

int direct_fd_of(PyObject *file) {
    PyObject *io = PyImport_ImportModule("io");
    if (!io) {
        PyErr_Clear();
        return -1;
    }
    PyObject *buffered = PyObject_GetAttrString(io, "BufferedReader"), *fileio = PyObject_GetAttrString(io, "FileIO");
    Py_DECREF(io);
    int fd = -1;
    if (buffered && fileio) {
        bool ok = false;
        if ((PyObject *)Py_TYPE(file) == fileio) ok = true;
        else if ((PyObject *)Py_TYPE(file) == buffered) {
            PyObject *raw = PyObject_GetAttrString(file, "raw");
            ok = raw && (PyObject *)Py_TYPE(raw) == fileio;
            Py_XDECREF(raw);
        }
        if (ok) {
            PyObject *r = PyObject_CallMethod(file, "fileno", nullptr);
            if (r) {
                const long v = PyLong_AsLong(r);
                Py_DECREF(r);
                struct stat st;
                if (v >= 0 && fstat((int)v, &st) == 0 && S_ISREG(st.st_mode)) fd = (int)v;
            }
        }
    }
    PyErr_Clear();
    Py_XDECREF(buffered);
    Py_XDECREF(fileio);
    return fd;
}

// bytes [pos, pos + len) of a regular file into dst from up to eight threads; what was read in one piece from
// `pos` on (short at the end of the file), or -1 with *err = errno
long long parallel_pread(int fd, uint8_t *dst, size_t len, long long pos, int *err) {
    unsigned n_threads = std::min<unsigned>(8, std::max(1u, std::thread::hardware_concurrency() / 2));
    n_threads = (unsigned)std::min<size_t>(n_threads, std::max<size_t>(1, len / DIRECT_MIN_BYTES));
    const size_t piece = ((len + n_threads - 1) / n_threads + 4095) & ~(size_t)4095;
    std::vector<long long> got(n_threads, 0);
    std::vector<int> errs(n_threads, 0);
    auto work = [&](unsigned k) {
        const size_t lo = std::min(len, k * piece), hi = std::min(len, lo + piece);
        size_t at = lo;
        while (at < hi) {
            const ssize_t r = pread(fd, dst + at, hi - at, (off_t)(pos + (long long)at));
            if (r < 0) {
                if (errno == EINTR) continue;
                errs[k] = errno;
                break;
            }
            if (r == 0) break;  // end of the file
            at += (size_t)r;
        }
        got[k] = (long long)(at - lo);
    };
    std::vector<std::thread> pool;
    for (unsigned k = 1; k < n_threads; k++) pool.emplace_back(work, k);
    work(0);
    for (auto &t : pool) t.join();
    long long total = 0;
    for (unsigned k = 0; k < n_threads; k++) {
        if (errs[k]) {
            *err = errs[k];
            return -1;
        }
        const size_t lo = std::min(len, k * piece), hi = std::min(len, lo + piece);
        total += got[k];
        if ((size_t)got[k] < hi - lo) break;  // the file ends inside this piece: later pieces lie behind the end
    }
    return total;
}

// fills dst[0..len) from the parser's file: the number of bytes read (0 at the end of the file), -1 with an exception
Py_ssize_t parser_read(Parser *self, uint8_t *dst, Py_ssize_t len) {
    if (self->direct_fd < 0 || (size_t)len < DIRECT_MIN_BYTES) return call_readinto(self->file, dst, len);
    PyObject *pos_o = PyObject_CallMethod(self->file, "tell", nullptr);
    if (!pos_o) return -1;
    const long long pos = PyLong_AsLongLong(pos_o);
    Py_DECREF(pos_o);
    if (pos < 0) {
        if (!PyErr_Occurred()) PyErr_SetString(PyExc_OSError, "negative file position");
        return -1;
    }
    int err = 0;
    long long total = 0;
    Py_BEGIN_ALLOW_THREADS
    total = parallel_pread(self->direct_fd, dst, (size_t)len, pos, &err);
    Py_END_ALLOW_THREADS
    if (total < 0) {
        errno = err;
        PyErr_SetFromErrno(PyExc_OSError);
        return -1;
    }
    PyObject *r = PyObject_CallMethod(self->file, "seek", "L", pos + total);
    if (!r) return -1;
    Py_DECREF(r);
    return (Py_ssize_t)total;
}
// ---- FastqParser iterating over a regular file: the file is read AHEAD of the parser -------------------------
// A third thread fetches the file block by block (parallel_pread) into pinned buffers while the read-ahead thread
// copies the previous block to the device and scans it and the caller's thread feeds the collectors with the one
// before: read, copy + scan and the collectors' kernels run side by side instead of one after the other.  A block
// is read RES bytes into its buffer, so that the unfinished record the previous block ended with is copied in
// front of it, not the block behind it.  The Python file object is not touched while the reader runs; when the
// reader ends (end of file, read(n), the parser's end) the object is moved behind the last block handed out.
struct FileReader {
    static constexpr size_t RES = (size_t)1 << 20;  // room in front of a block for the previous block's leftover
    static constexpr int SLOTS = 2;  // one block being read, one waiting for the parser
    struct Slot {
        Pinned buf;      // RES + step bytes, allocated by the consumer (the pinned pool belongs to the GIL holder)
        size_t got = 0;
        bool full = false;
    };
    int fd = -1;
    size_t step = 0;
    long long next_off = 0;      // where the worker reads next
    long long consumed_off = 0;  // the file position behind the last block handed out
    Slot slots[SLOTS];
    unsigned head = 0, tail = 0;  // next slot to hand out / to fill
    bool eof = false, stop = false;
    int err = 0;
    std::mutex mu;
    std::condition_variable cv;
    std::thread worker;

    void run() {
        for (;;) {
            Slot *s;
            long long off;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || (!slots[tail % SLOTS].full && slots[tail % SLOTS].buf.ptr && tail - head < SLOTS); });
                if (stop) return;
                s = &slots[tail % SLOTS];
                off = next_off;
            }
            int e = 0;
            const long long got = parallel_pread(fd, s->buf.ptr + RES, step, off, &e);
            std::lock_guard<std::mutex> lk(mu);
            if (got < 0) {
                err = e;
                eof = true;
            }
            else {
                s->got = (size_t)got;
                s->full = true;
                next_off += got;
                tail++;
                if ((size_t)got < step) eof = true;
            }
            cv.notify_all();
            if (eof) return;
        }
    }
};

// stops the reader and moves the file object behind what the parser has been given (seek: with the GIL held)
void reader_stop(Parser *self, bool seek) {
    FileReader *r = self->reader;
    if (!r) return;
    self->reader = nullptr;
    self->reader_done = true;
    {
        std::lock_guard<std::mutex> lk(r->mu);
        r->stop = true;
    }
    r->cv.notify_all();
    Py_BEGIN_ALLOW_THREADS
    if (r->worker.joinable()) r->worker.join();
    Py_END_ALLOW_THREADS
    for (auto &s : r->slots) s.buf.release();
    if (seek) {
        PyObject *t = nullptr, *v = nullptr, *tb = nullptr;
        PyErr_Fetch(&t, &v, &tb);  // (may run while an exception travels)
        PyObject *res = PyObject_CallMethod(self->file, "seek", "L", r->consumed_off);
        if (!res) PyErr_Clear();  // best effort: a closed file cannot be moved, and nobody will read it either
        Py_XDECREF(res);
        PyErr_Restore(t, v, tb);
    }
    delete r;
}

// first call: reader from the file object's position.  false with an exception set, or with none when this file
// is read the classic way
bool reader_start(Parser *self) {
    PyObject *pos_o = PyObject_CallMethod(self->file, "tell", nullptr);
    if (!pos_o) return false;
    const long long pos = PyLong_AsLongLong(pos_o);
    Py_DECREF(pos_o);
    if (pos < 0) {
        PyErr_Clear();
        self->reader_done = true;
        return false;
    }
    FileReader *r = new FileReader();
    r->fd = self->direct_fd;
    r->step = (size_t)self->read_in_size;
    r->next_off = r->consumed_off = pos;
    for (auto &s : r->slots)
        if (!s.buf.alloc(FileReader::RES + r->step)) {
            for (auto &t : r->slots) t.buf.release();
            delete r;
            return false;
        }
    r->worker = std::thread([r] { r->run(); });
    self->reader = r;
    return true;
}

// the next block, with `leftover` copied in front of it: text = buf.ptr + off, n bytes, `last` when the file ends
// with it.  0 on success, -1 with an exception, 1 when the reader has nothing more (it is stopped then).
int reader_take(Parser *self, const std::vector<uint8_t> &leftover, Pinned *buf, size_t *off, size_t *n, bool *last) {
    FileReader *r = self->reader;
    FileReader::Slot *s = nullptr;
    bool drained = false;
    int err = 0;
    Py_BEGIN_ALLOW_THREADS
    {
        std::unique_lock<std::mutex> lk(r->mu);
        r->cv.wait(lk, [&] { return r->slots[r->head % FileReader::SLOTS].full || (r->eof && r->head == r->tail); });
        if (r->slots[r->head % FileReader::SLOTS].full) s = &r->slots[r->head % FileReader::SLOTS];
        else drained = true;
        err = r->err;
    }
    Py_END_ALLOW_THREADS
    if (!s || drained) {
        reader_stop(self, true);
        if (err) {
            errno = err;
            PyErr_SetFromErrno(PyExc_OSError);
            return -1;
        }
        return 1;
    }
    if (leftover.size() > FileReader::RES) {  // a record longer than RES: its block moves behind it in a buffer of its own
        Pinned big;
        if (!big.alloc(leftover.size() + s->got)) return -1;
        memcpy(big.ptr, leftover.data(), leftover.size());
        memcpy(big.ptr + leftover.size(), s->buf.ptr + FileReader::RES, s->got);
        *buf = big;
        *off = 0;
        *n = leftover.size() + s->got;
        *last = s->got < r->step;
        std::lock_guard<std::mutex> lk(r->mu);  // the slot keeps its buffer
        r->consumed_off += (long long)s->got;
        s->full = false;
        r->head++;
        r->cv.notify_all();
        return 0;
    }
    Pinned fresh;  // the block's buffer goes with the record array; the slot gets a new one
    const bool more = s->got == r->step;
    if (more && !fresh.alloc(FileReader::RES + r->step)) return -1;
    if (!leftover.empty()) memcpy(s->buf.ptr + FileReader::RES - leftover.size(), leftover.data(), leftover.size());
    *buf = s->buf;
    *off = FileReader::RES - leftover.size();
    *n = leftover.size() + s->got;
    *last = !more;
    std::lock_guard<std::mutex> lk(r->mu);
    r->consumed_off += (long long)s->got;
    s->buf = fresh;
    s->full = false;
    r->head++;
    r->cv.notify_all();
    return 0;
}

PyObject *raise_format_error(const uint8_t *data, uint64_t nbytes, const sq_parse_info &info) {
    const uint64_t pos = info.err_pos < nbytes ? info.err_pos : 0;
    switch (info.err_code) {
    case SQ_PARSE_ASCII: {
        PyObject *c = PyUnicode_DecodeLatin1((const char *)data + pos, 1, nullptr);
        if (c) PyErr_Format(PyExc_ValueError, "Found non-ASCII character in file: %U", c);
        Py_XDECREF(c);
        return nullptr;
    }
    case SQ_PARSE_NO_AT:
        PyErr_Format(PyExc_ValueError, "Record does not start with @ but with %c", (int)data[pos]);
        return nullptr;
    case SQ_PARSE_NO_PLUS:
        PyErr_Format(PyExc_ValueError, "Record second header does not start with + but with %c", (int)data[pos]);
        return nullptr;
    default: {
        uint64_t end = pos;
        while (end < nbytes && data[end] != '\n') end++;
        PyObject *name = PyUnicode_DecodeASCII((const char *)data + pos, (Py_ssize_t)(end - pos), "replace");
        if (name) PyErr_Format(PyExc_ValueError, "Record sequence and qualities do not have equal length, %R", name);
        Py_XDECREF(name);
        return nullptr;
    }
    }
}
// iteration over a regular file: blocks from the FileReader.  nullptr + no exception: go on the classic way
PyObject *FQ_from_reader(Parser *self) {
    sq_ctx *ctx = g_ctx;
    std::vector<uint8_t> &leftover = *self->leftover;
    for (;;) {
        Pinned buf;
        size_t off = 0, n = 0;
        bool last = false;
        const int st = reader_take(self, leftover, &buf, &off, &n, &last);
        if (st != 0) return nullptr;  // an exception (-1), or the end of the file (1: the classic path sees it too)
        if (last) {
            // the last block: it becomes the leftover and the classic path finishes (it reads the end of the file
            // and words the errors of an unfinished record)
            leftover.assign(buf.ptr + off, buf.ptr + off + n);
            buf.release();
            reader_stop(self, true);
            return nullptr;
        }
        sq_batch *handle = nullptr;
        sq_parse_info info;
        memset(&info, 0, sizeof(info));
        int rc;
        Py_BEGIN_ALLOW_THREADS
        rc = sq_batch_from_fastq(ctx, buf.ptr + off, n, UINT64_MAX, &handle, &info);
        Py_END_ALLOW_THREADS
        if (rc != SQ_OK) {
            if (rc == SQ_E_FORMAT) raise_format_error(buf.ptr + off, n, info);
            else raise_sq(rc, "sq_batch_from_fastq");
            buf.release();
            reader_stop(self, true);
            return nullptr;
        }
        if (info.n_records == 0) {  // no whole record yet (a record longer than a block): on with the next block behind it
            if (handle) sq_batch_free(handle);
            leftover.assign(buf.ptr + off, buf.ptr + off + n);
            buf.release();
            continue;
        }
        leftover.assign(buf.ptr + off + info.consumed, buf.ptr + off + n);
        ArrayView *a = ArrayView_alloc();
        if (!a) {
            sq_batch_free(handle);
            buf.release();
            return nullptr;
        }
        a->h = handle;
        a->n = info.n_records;
        a->nbytes = n;
        a->pinned = buf;
        a->pinned_off = off;
        return (PyObject *)a;
    }
}

PyObject *FQ_create_record_array(Parser *self, uint64_t min_records, uint64_t max_records) {
    sq_ctx *ctx = g_ctx;
    if (min_records == 1 && max_records == UINT64_MAX && self->direct_fd >= 0 && !self->reader_done &&
        (size_t)self->read_in_size >= DIRECT_MIN_BYTES) {
        if (!self->reader && !reader_start(self) && PyErr_Occurred()) return nullptr;
        if (self->reader) {
            PyObject *a = FQ_from_reader(self);
            if (a || PyErr_Occurred()) return a;
        }
    }
    std::vector<uint8_t> &leftover = *self->leftover;
    const size_t step = (size_t)self->read_in_size;
    const size_t size = std::max(step, leftover.size() + (leftover.size() < step ? 0 : step));
    Pinned buf;
    if (!buf.alloc(size)) return nullptr;
    if (!leftover.empty()) memcpy(buf.ptr, leftover.data(), leftover.size());
    size_t filled = leftover.size();
    uint64_t parsed = 0;
    sq_batch *handle = nullptr;
    sq_parse_info info;
    memset(&info, 0, sizeof(info));
    auto fail = [&]() -> PyObject * {
        if (handle) sq_batch_free(handle);
        buf.release();
        return nullptr;
    };
    while (parsed < min_records) {
        if (filled == buf.size) {  // grow by one step, keeping the content (:995-1020)
            Pinned bigger;
            if (!bigger.alloc(buf.size + step)) return fail();
            memcpy(bigger.ptr, buf.ptr, filled);
            buf.release();
            buf = bigger;
        }
        const Py_ssize_t got = parser_read(self, buf.ptr + filled, (Py_ssize_t)(buf.size - filled));
        if (got < 0) return fail();
        const size_t new_filled = filled + (size_t)got;
        if (new_filled == 0) break;  // entire file is read
        if (handle) {
            sq_batch_free(handle);
            handle = nullptr;
        }
        int rc;
        Py_BEGIN_ALLOW_THREADS
        rc = sq_batch_from_fastq(ctx, buf.ptr, new_filled, max_records, &handle, &info);
        Py_END_ALLOW_THREADS
        if (rc == SQ_E_FORMAT) {
            handle = nullptr;
            raise_format_error(buf.ptr, new_filled, info);
            return fail();
        }
        if (rc != SQ_OK) {
            handle = nullptr;
            raise_sq(rc, "sq_batch_from_fastq");
            return fail();
        }
        parsed = info.n_records;
        if (got == 0) {
            size_t newlines = 0;
            for (size_t i = 0; i < new_filled && newlines < 4; i++) newlines += buf.ptr[i] == '\n';
            if (newlines < 4) {
                size_t end = 0;
                while (end < new_filled && buf.ptr[end] != 0) end++;
                PyObject *text = PyUnicode_DecodeASCII((const char *)buf.ptr, (Py_ssize_t)end, "replace");
                if (text) PyErr_Format(PyExc_EOFError, "Incomplete record at the end of file %U", text);
                Py_XDECREF(text);
                return fail();
            }
            filled = new_filled;
            break;
        }
        filled = new_filled;
    }
    if (!handle) {
        leftover.clear();
        buf.release();
        return ArrayView_empty();
    }
    leftover.assign(buf.ptr + info.consumed, buf.ptr + filled);
    ArrayView *a = ArrayView_alloc();
    if (!a) return fail();
    a->h = handle;
    a->n = parsed;
    a->nbytes = filled;
    a->pinned = buf;
    return (PyObject *)a;
}
PyObject *FQ_produce(Parser *self) { return FQ_create_record_array(self, 1, UINT64_MAX); }
PyObject *FQ_next(Parser *self) {
    PyObject *a = self->ra->active ? read_ahead_take(self) : FQ_produce(self);
    if (a && ((ArrayView *)a)->n == 0) {
        Py_DECREF(a);
        return nullptr;  // StopIteration
    }
    if (a && self->read_in_size >= READ_AHEAD_MIN_STEP) read_ahead_start(self, FQ_produce);
    return a;
}
PyObject *FQ_read(Parser *self, PyObject *n_o) {
    const Py_ssize_t n = PyNumber_AsSsize_t(n_o, PyExc_OverflowError);
    if (n == -1 && PyErr_Occurred()) return nullptr;
    if (n < 1) {
        PyErr_Format(PyExc_ValueError, "number_of_records should be greater than 1, got %zd", n);
        return nullptr;
    }
    if (self->ra->active) {
        // an array was read ahead by __next__: un-read it (its buffer = the leftover it started from + what it read)
        PyObject *ahead = read_ahead_take(self);
        if (!ahead) return nullptr;
        ArrayView *av = (ArrayView *)ahead;
        if (av->pinned.ptr) self->leftover->assign(av->pinned.ptr + av->pinned_off, av->pinned.ptr + av->pinned_off + av->nbytes);
        Py_DECREF(ahead);
    }
    reader_stop(self, true);  // (blocks read ahead of the parser are dropped: the file object goes back behind the last one used)
    return FQ_create_record_array(self, (uint64_t)n, (uint64_t)n);
}
PyMethodDef FQ_methods[] = {{"read", (PyCFunction)FQ_read, METH_O, "Read up to number_of_records records."},
                            {nullptr, nullptr, 0, nullptr}};
PyType_Slot FQ_slots[] = {{Py_tp_new, (void *)FQ_new},           {Py_tp_dealloc, (void *)Parser_dealloc},
                          {Py_tp_iter, (void *)Parser_iter},     {Py_tp_iternext, (void *)FQ_next},
                          {Py_tp_methods, FQ_methods},           {0, nullptr}};
PyType_Spec FQ_spec = {"_qc.FastqParser", sizeof(Parser), 0, Py_TPFLAGS_DEFAULT, FQ_slots};

// ---------------------------------------------------------------------------------------------
// BamParser (reference :1362-1725): header skip here; the block_size chain, the flag drop and the 4-bit sequence /
// raw quality decode on the device (sq_batch_from_bam_bytes)
// ---------------------------------------------------------------------------------------------
PyObject *read_exact(PyObject *file, Py_ssize_t n, bool first) {
    PyObject *b = PyObject_CallMethod(file, "read", "n", n);
    if (!b) return nullptr;
    if (!PyBytes_CheckExact(b)) {
        if (first) PyErr_Format(PyExc_TypeError, "file_obj %R is not a binary IO type, got %R", file, (PyObject *)Py_TYPE(file));
        else PyErr_SetString(PyExc_TypeError, "read() did not return bytes");
        Py_DECREF(b);
        return nullptr;
    }
    if (PyBytes_GET_SIZE(b) != n) {
        Py_DECREF(b);
        PyErr_SetString(PyExc_EOFError, "Truncated BAM file");
        return nullptr;
    }
    return b;
}
uint32_t le32(const void *p) {
    const uint8_t *b = (const uint8_t *)p;
    return (uint32_t)b[0] | (uint32_t)b[1] << 8 | (uint32_t)b[2] << 16 | (uint32_t)b[3] << 24;
}
PyObject *BAM_new(PyTypeObject *type, PyObject *args, PyObject *kwargs) {
    static const char *kw[] = {"fileobj", "initial_buffersize", nullptr};
    PyObject *file = nullptr, *size_o = Py_None;
    if (!PyArg_ParseTupleAndKeywords(args, kwargs, "O|O:BamParser", (char **)kw, &file, &size_o)) return nullptr;
    Py_ssize_t size = DEFAULT_BAM_BUFFERSIZE;
    if (size_o != Py_None) {
        size = PyLong_AsSsize_t(size_o);
        if (size == -1 && PyErr_Occurred()) return nullptr;
    }
    if (size < 4) {
        PyErr_Format(PyExc_ValueError, "initial_buffersize must be at least 4, got %zd", size);
        return nullptr;
    }
    PyObject *magic = read_exact(file, 8, true);
    if (!magic) return nullptr;
    if (memcmp(PyBytes_AS_STRING(magic), "BAM\1", 4) != 0) {
        PyErr_Format(PyExc_ValueError, "fileobj: %R, is not a BAM file. No BAM magic, instead found: %R", file, magic);
        Py_DECREF(magic);
        return nullptr;
    }
    const uint32_t l_text = le32(PyBytes_AS_STRING(magic) + 4);
    Py_DECREF(magic);
    PyObject *header = read_exact(file, (Py_ssize_t)l_text, false);
    if (!header) return nullptr;
    PyObject *n_ref_b = read_exact(file, 4, false);
    if (!n_ref_b) {
        Py_DECREF(header);
        return nullptr;
    }
    const uint32_t n_ref = le32(PyBytes_AS_STRING(n_ref_b));
    Py_DECREF(n_ref_b);
    for (uint32_t i = 0; i < n_ref; i++) {
        PyObject *ln = read_exact(file, 4, false);
        PyObject *rest = ln ? read_exact(file, (Py_ssize_t)le32(PyBytes_AS_STRING(ln)) + 4, false) : nullptr;
        Py_XDECREF(ln);
        if (!rest) {
            Py_DECREF(header);
            return nullptr;
        }
        Py_DECREF(rest);
    }
    Parser *self = ctx_get() ? (Parser *)type->tp_alloc(type, 0) : nullptr;
    if (!self) {
        Py_DECREF(header);
        return nullptr;
    }
    Py_INCREF(file);
    self->file = file;
    self->header = header;
    self->read_in_size = size;
    self->leftover = new std::vector<uint8_t>();
    new (&self->bam_buf) Pinned();
    self->bam_filled = 0;
    self->bam_n_ref = (int32_t)std::min<uint32_t>(n_ref, 0x7fffffffu);
    self->direct_fd = size >= (Py_ssize_t)READ_AHEAD_MIN_STEP ? direct_fd_of(file) : -1;
    self->reader = nullptr;
    self->reader_done = false;
    self->ra = new ReadAhead();
    return (PyObject *)self;
}
PyObject *BAM_produce(Parser *self);
PyObject *BAM_next(Parser *self) {
    PyObject *a = self->ra->active ? read_ahead_take(self) : BAM_produce(self);
    if (a && self->read_in_size >= READ_AHEAD_MIN_STEP) read_ahead_start(self, BAM_produce);
    return a;
}
// iteration over a regular file: blocks from the FileReader (see FQ_from_reader).  nullptr + no exception: the
// classic path below goes on (with what is left of the last block in bam_buf)
PyObject *BAM_from_reader(Parser *self) {
    std::vector<uint8_t> &leftover = *self->leftover;
    for (;;) {
        Pinned buf;
        size_t off = 0, n = 0;
        bool last = false;
        const int st = reader_take(self, leftover, &buf, &off, &n, &last);
        if (st != 0) return nullptr;
        sq_batch *h = nullptr;
        uint64_t kept = 0, skipped = 0, consumed = 0, packed = 0;
        int rc;
        Py_BEGIN_ALLOW_THREADS
        rc = sq_batch_from_bam_bytes(g_ctx, buf.ptr + off, n, self->bam_n_ref, &h, &kept, &skipped, &consumed, &packed);
        Py_END_ALLOW_THREADS
        if (rc != SQ_OK) {
            buf.release();
            reader_stop(self, true);
            return raise_sq(rc, "sq_batch_from_bam_bytes");
        }
        leftover.assign(buf.ptr + off + consumed, buf.ptr + off + n);
        buf.release();
        if (last) {  // the classic path ends the file: nothing left -> StopIteration, an unfinished record -> EOFError
            Pinned &bb = self->bam_buf;
            if (bb.size < leftover.size()) {
                bb.release();
                if (!bb.alloc(leftover.size())) {
                    if (h) sq_batch_free(h);
                    return nullptr;
                }
            }
            if (!leftover.empty()) memcpy(bb.ptr, leftover.data(), leftover.size());
            self->bam_filled = leftover.size();
            leftover.clear();
            reader_stop(self, true);
        }
        if (!kept && !skipped && !last) continue;  // no whole record yet: on with the next block behind it
        if (!kept) {
            if (h) sq_batch_free(h);
            return last && !skipped ? nullptr : ArrayView_empty();
        }
        ArrayView *a = ArrayView_alloc();
        if (!a) {
            sq_batch_free(h);
            return nullptr;
        }
        a->h = h;
        a->n = kept;
        a->nbytes = packed;
        return (PyObject *)a;
    }
}

PyObject *BAM_produce(Parser *self) {
    if (self->direct_fd >= 0 && !self->reader_done && (size_t)self->read_in_size >= DIRECT_MIN_BYTES && self->bam_filled == 0) {
        if (!self->reader && !reader_start(self) && PyErr_Occurred()) return nullptr;
        if (self->reader) {
            PyObject *a = BAM_from_reader(self);
            if (a || PyErr_Occurred()) return a;
        }
    }
    // [leftover | newly read bytes] live in one pinned buffer that is kept between calls: the
    // host->device copy of sq_batch_from_bam_bytes runs at PCIe speed and nothing is re-allocated or zeroed
    Pinned &buf = self->bam_buf;
    const size_t step = (size_t)self->read_in_size;
    uint64_t consumed = 0, kept = 0, skipped = 0, packed = 0;
    sq_batch *h = nullptr;
    for (;;) {
        const size_t have = self->bam_filled;
        const size_t want = have >= 4 ? std::max<size_t>(le32(buf.ptr), step) : step - have;  // :1527-1531
        if (have + want > buf.size) {
            Pinned bigger;
            if (!bigger.alloc(std::max(have + want, buf.size + buf.size / 2))) return nullptr;
            if (have) memcpy(bigger.ptr, buf.ptr, have);
            buf.release();
            buf = bigger;
        }
        const Py_ssize_t got = parser_read(self, buf.ptr + have, (Py_ssize_t)want);
        if (got < 0) return nullptr;
        const size_t n = have + (size_t)got;
        if (n == 0) return nullptr;  // StopIteration
        if (got == 0) {
            PyObject *b = PyBytes_FromStringAndSize((const char *)buf.ptr, (Py_ssize_t)have);
            if (b) PyErr_Format(PyExc_EOFError, "Incomplete record at the end of file %R", b);
            Py_XDECREF(b);
            return nullptr;
        }
        self->bam_filled = n;
        // record chain (:1623-1637), flag drop and decode of the bytes read so far: on the device
        int rc;
        Py_BEGIN_ALLOW_THREADS
        rc = sq_batch_from_bam_bytes(g_ctx, buf.ptr, n, self->bam_n_ref, &h, &kept, &skipped, &consumed, &packed);  // (synchronises)
        Py_END_ALLOW_THREADS
        if (rc != SQ_OK) return raise_sq(rc, "sq_batch_from_bam_bytes");
        if (kept || skipped) break;
    }
    ArrayView *a = nullptr;
    if (kept) {
        a = ArrayView_alloc();
        if (!a) {
            sq_batch_free(h);
            return nullptr;
        }
        a->h = h;
        a->n = kept;
        a->nbytes = packed;
    }
    const size_t left = self->bam_filled - (size_t)consumed;
    if (left && consumed) memmove(buf.ptr, buf.ptr + consumed, left);
    self->bam_filled = left;
    return a ? (PyObject *)a : ArrayView_empty();
}
PyMemberDef BAM_members[] = {{"header", T_OBJECT, offsetof(Parser, header), READONLY, "the SAM header text"},
                             {nullptr, 0, 0, 0, nullptr}};
PyType_Slot BAM_slots[] = {{Py_tp_new, (void *)BAM_new},         {Py_tp_dealloc, (void *)Parser_dealloc},
                           {Py_tp_iter, (void *)Parser_iter},    {Py_tp_iternext, (void *)BAM_next},
                           {Py_tp_members, BAM_members},         {0, nullptr}};
PyType_Spec BAM_spec = {"_qc.BamParser", sizeof(Parser), 0, Py_TPFLAGS_DEFAULT, BAM_slots};

// ---------------------------------------------------------------------------------------------
// module
// ---------------------------------------------------------------------------------------------
PyObject *mod_sync(PyObject *, PyObject *) {
    if (flush_pending() < 0) return nullptr;
    if (g_ctx) SQ_CHECK(sq_ctx_sync(g_ctx), "sq_ctx_sync");
    Py_RETURN_NONE;
}
PyObject *mod_launch_count(PyObject *, PyObject *) { return PyLong_FromUnsignedLongLong(g_ctx ? sq_ctx_launch_count(g_ctx) : 0); }
PyMethodDef module_methods[] = {
    {"_sync", mod_sync, METH_NOARGS, "apply the pending adds and wait for the device (extension of the B200 build)"},
    {"_launch_count", mod_launch_count, METH_NOARGS, "kernels launched so far (extension of the B200 build)"},
    {nullptr, nullptr, 0, nullptr}};

struct TypeEntry { const char *name; PyType_Spec *spec; PyTypeObject **slot; };
PyTypeObject *QCType, *ADType, *PTType, *OVType, *DDType, *NSType, *ISType, *FQType, *BAMType;

int module_exec(PyObject *m) {
    SqGpuError = PyErr_NewException("_qc.SqGpuError", PyExc_RuntimeError, nullptr);
    if (!SqGpuError || PyModule_AddObject(m, "SqGpuError", SqGpuError) < 0) return -1;
    Py_INCREF(SqGpuError);
    PyObject *array_mod = PyImport_ImportModule("array");
    if (!array_mod) return -1;
    g_array_type = PyObject_GetAttrString(array_mod, "array");
    Py_DECREF(array_mod);
    if (!g_array_type) return -1;
    TypeEntry types[] = {{"FastqRecordView", &RecordView_spec, &RecordViewType},
                         {"FastqRecordArrayView", &ArrayView_spec, &ArrayViewType},
                         {"FastqParser", &FQ_spec, &FQType},
                         {"BamParser", &BAM_spec, &BAMType},
                         {"QCMetrics", &QC_spec, &QCType},
                         {"AdapterCounter", &AD_spec, &ADType},
                         {"PerTileQuality", &PT_spec, &PTType},
                         {"OverrepresentedSequences", &OV_spec, &OVType},
                         {"DedupEstimator", &DD_spec, &DDType},
                         {"NanoStats", &NS_spec, &NSType},
                         {"NanoporeReadInfo", &ReadInfo_spec, &ReadInfoType},
                         {"NanoStatsIterator", &NanoIter_spec, &NanoIterType},
                         {"InsertSizeMetrics", &IS_spec, &ISType}};
    for (auto &t : types) {
        PyObject *tp = PyType_FromSpec(t.spec);
        if (!tp) return -1;
        *t.slot = (PyTypeObject *)tp;
        Py_INCREF(tp);
        if (PyModule_AddObject(m, t.name, tp) < 0) return -1;
    }
    // module constants (reference _qcmodule.c:6082-6171)
    struct { const char *name; long long v; } consts[] = {
        {"A", 0}, {"C", 1}, {"G", 2}, {"T", 3}, {"N", 4},
        {"NUMBER_OF_NUCS", 5}, {"NUMBER_OF_PHREDS", 12}, {"TABLE_SIZE", 60}, {"PHRED_MAX", 93},
        {"MAX_SEQUENCE_SIZE", 64}, {"DEFAULT_END_ANCHOR_LENGTH", 100},
        {"DEFAULT_MAX_UNIQUE_FRAGMENTS", 5000000}, {"DEFAULT_DEDUP_MAX_STORED_FINGERPRINTS", 1000000},
        {"DEFAULT_FRAGMENT_LENGTH", 21}, {"DEFAULT_UNIQUE_SAMPLE_EVERY", 8},
        {"DEFAULT_BASES_FROM_START", 100}, {"DEFAULT_BASES_FROM_END", 100},
        {"DEFAULT_FINGERPRINT_FRONT_SEQUENCE_LENGTH", 8}, {"DEFAULT_FINGERPRINT_BACK_SEQUENCE_LENGTH", 8},
        {"DEFAULT_FINGERPRINT_FRONT_SEQUENCE_OFFSET", 64}, {"DEFAULT_FINGERPRINT_BACK_SEQUENCE_OFFSET", 64},
        {"INSERT_SIZE_MAX_ADAPTER_STORE_SIZE", 31}};
    for (auto &c : consts)
        if (PyModule_AddIntConstant(m, c.name, (long)c.v) < 0) return -1;
    return 0;
}

PyModuleDef_Slot module_slots[] = {{Py_mod_exec, (void *)module_exec}, {0, nullptr}};
PyModuleDef module_def = {PyModuleDef_HEAD_INIT, "_qc",
                          "B200 (sm_100a) build of sequali's native QC module: same API, kernels in libsqgpu.so", 0,
                          module_methods, module_slots, nullptr, nullptr, nullptr};

}  // namespace

PyMODINIT_FUNC PyInit__qc(void) { return PyModuleDef_Init(&module_def); }

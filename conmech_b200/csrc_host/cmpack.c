/* _cmpack - packs a Python sequence of element-id sequences into subset bit masks
 * (uint32 words, bit e%32 of word e/32 = element e exists), the wire format of cm_solve_many.
 * The batched solve consumes several hundred thousand subsets per second and GPU; converting
 * Python lists through numpy costs ~70 ns per id, this loop 2.5-4 ns (ints read straight from the objects).  An empty sequence means the
 * full structure (reference numpy_stiffness.py:230-231); an id outside [0, nE) raises IndexError
 * with the reference's message (stiffness_checker.py:178-181); a repeated id raises ValueError.
 * publish(ready, k, value): release-store into an int32 flag array (the hand-over of a packed chunk to
 * the solving thread, cm_solve_many_host_chunks).
 *
 *   _cmpack.pack(list_of_ids, nE, out)      out: writable C-contiguous buffer of B*ceil(nE/32) uint32
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <string.h>

static int set_bit(uint32_t *row, Py_ssize_t e, Py_ssize_t nE) {
    if (e < 0 || e >= nE) {
        PyErr_SetString(PyExc_IndexError, "element id not within range!");
        return -1;
    }
    const uint32_t bit = (uint32_t)1u << (e & 31);
    if (row[e >> 5] & bit) {
        /* The reference assembles a repeated id twice (numpy_stiffness_assembly.py:371-388 has no
         * de-duplication: the element's stiffness is double counted).  A bit mask is a set and cannot
         * carry that; rather than silently solving a different structure, reject the input. */
        PyErr_Format(PyExc_ValueError, "duplicate element id %zd in a subset (the reference would double count it)", e);
        return -1;
    }
    row[e >> 5] |= bit;
    return 0;
}

/* One element id from a Python object.  Exact ints that fit one digit (every id below 2^30: all of them in practice)
 * are read straight from the object (CPython >= 3.12 "compact" ints); everything else - numpy scalars, objects with
 * __index__, big ints - goes through the number protocol.  -1 with an exception set on failure. */
static inline Py_ssize_t id_from_object(PyObject *o) {
#if PY_VERSION_HEX >= 0x030C0000
    if (PyLong_CheckExact(o) && PyUnstable_Long_IsCompact((PyLongObject *)o))
        return PyUnstable_Long_CompactValue((PyLongObject *)o);
#endif
    return PyNumber_AsSsize_t(o, PyExc_IndexError);
}

static void set_all(uint32_t *row, Py_ssize_t nE) {
    Py_ssize_t full = nE >> 5;
    for (Py_ssize_t w = 0; w < full; ++w) row[w] = 0xffffffffu;
    if (nE & 31) row[full] = ((uint32_t)1u << (nE & 31)) - 1u;
}

static PyObject *pack(PyObject *self, PyObject *args) {
    PyObject *subsets, *outobj;
    Py_ssize_t nE;
    if (!PyArg_ParseTuple(args, "OnO", &subsets, &nE, &outobj)) return NULL;
    if (nE <= 0) { PyErr_SetString(PyExc_ValueError, "nE must be positive"); return NULL; }
    PyObject *outer = PySequence_Fast(subsets, "pack: expected a sequence of id sequences");
    if (!outer) return NULL;
    const Py_ssize_t B = PySequence_Fast_GET_SIZE(outer);
    const Py_ssize_t W = (nE + 31) / 32;
    Py_buffer ob;
    if (PyObject_GetBuffer(outobj, &ob, PyBUF_WRITABLE | PyBUF_C_CONTIGUOUS) < 0) { Py_DECREF(outer); return NULL; }
    if (ob.len != (Py_ssize_t)(B * W * 4)) {
        PyBuffer_Release(&ob); Py_DECREF(outer);
        PyErr_SetString(PyExc_ValueError, "pack: output buffer must hold B*ceil(nE/32) uint32 words");
        return NULL;
    }
    uint32_t *out = (uint32_t *)ob.buf;
    memset(out, 0, (size_t)ob.len);
    int fail = 0;
    for (Py_ssize_t b = 0; b < B && !fail; ++b) {
        uint32_t *row = out + b * W;
        PyObject *item = PySequence_Fast_GET_ITEM(outer, b);
        if (PyList_Check(item) || PyTuple_Check(item)) {
            const Py_ssize_t n = PySequence_Fast_GET_SIZE(item);
            PyObject **it = PySequence_Fast_ITEMS(item);
            if (n == 0) { set_all(row, nE); continue; }
            for (Py_ssize_t i = 0; i < n; ++i) {
                if (i + 16 < n) __builtin_prefetch(it[i + 16]);     /* the int objects are what misses the cache */
                Py_ssize_t e = id_from_object(it[i]);
                if (e == -1 && PyErr_Occurred()) { fail = 1; break; }
                if (set_bit(row, e, nE) < 0) { fail = 1; break; }
            }
            continue;
        }
        /* integer arrays through the buffer protocol (numpy int64 / int32, C-contiguous) */
        Py_buffer ib;
        if (PyObject_CheckBuffer(item) && PyObject_GetBuffer(item, &ib, PyBUF_FORMAT | PyBUF_C_CONTIGUOUS) == 0) {
            const char *f = ib.format ? ib.format : "B";
            if (*f == '<' || *f == '=' || *f == '@') ++f;
            const Py_ssize_t n = ib.itemsize ? ib.len / ib.itemsize : 0;
            int handled = 1;
            if (n == 0) set_all(row, nE);
            else if ((*f == 'l' || *f == 'q') && ib.itemsize == 8) {
                const int64_t *p = (const int64_t *)ib.buf;
                for (Py_ssize_t i = 0; i < n; ++i) if (set_bit(row, (Py_ssize_t)p[i], nE) < 0) { fail = 1; break; }
            } else if ((*f == 'i' || *f == 'l') && ib.itemsize == 4) {
                const int32_t *p = (const int32_t *)ib.buf;
                for (Py_ssize_t i = 0; i < n; ++i) if (set_bit(row, (Py_ssize_t)p[i], nE) < 0) { fail = 1; break; }
            } else handled = 0;
            PyBuffer_Release(&ib);
            if (handled) continue;
        } else PyErr_Clear();
        /* anything else iterable (sets, ranges, generators, arrays of other integer types) */
        PyObject *seq = PySequence_Fast(item, "pack: every subset must be a sequence of element ids");
        if (!seq) { fail = 1; break; }
        const Py_ssize_t n = PySequence_Fast_GET_SIZE(seq);
        PyObject **it = PySequence_Fast_ITEMS(seq);
        if (n == 0) set_all(row, nE);
        for (Py_ssize_t i = 0; i < n; ++i) {
            Py_ssize_t e = id_from_object(it[i]);
            if ((e == -1 && PyErr_Occurred()) || set_bit(row, e, nE) < 0) { fail = 1; break; }
        }
        Py_DECREF(seq);
    }
    PyBuffer_Release(&ob);
    Py_DECREF(outer);
    if (fail) return NULL;
    Py_RETURN_NONE;
}

/* pack_csr(flat_ids, offsets, nE, out): subset b = flat_ids[offsets[b] : offsets[b+1]] (int64 buffers),
 * the form a planner that keeps its subsets in arrays already has - no Python objects per id. */
static PyObject *pack_csr(PyObject *self, PyObject *args) {
    PyObject *flatobj, *offobj, *outobj;
    Py_ssize_t nE;
    if (!PyArg_ParseTuple(args, "OOnO", &flatobj, &offobj, &nE, &outobj)) return NULL;
    if (nE <= 0) { PyErr_SetString(PyExc_ValueError, "nE must be positive"); return NULL; }
    Py_buffer fb, ofb, ob;
    if (PyObject_GetBuffer(flatobj, &fb, PyBUF_C_CONTIGUOUS | PyBUF_FORMAT) < 0) return NULL;
    if (PyObject_GetBuffer(offobj, &ofb, PyBUF_C_CONTIGUOUS | PyBUF_FORMAT) < 0) { PyBuffer_Release(&fb); return NULL; }
    if (PyObject_GetBuffer(outobj, &ob, PyBUF_WRITABLE | PyBUF_C_CONTIGUOUS) < 0) { PyBuffer_Release(&fb); PyBuffer_Release(&ofb); return NULL; }
    int ok = 0;
    const Py_ssize_t W = (nE + 31) / 32;
    const Py_ssize_t B = ofb.itemsize == 8 ? ofb.len / 8 - 1 : -1;
    if (fb.itemsize != 8 || ofb.itemsize != 8 || B < 0 || ob.len != (Py_ssize_t)(B * W * 4)) {
        PyErr_SetString(PyExc_ValueError, "pack_csr: int64 flat ids, int64 offsets [B+1], output of B*ceil(nE/32) uint32 words");
    } else {
        const int64_t *flat = (const int64_t *)fb.buf, *off = (const int64_t *)ofb.buf;
        const int64_t n_flat = fb.len / 8;
        uint32_t *out = (uint32_t *)ob.buf;
        memset(out, 0, (size_t)ob.len);
        ok = 1;
        for (Py_ssize_t b = 0; b < B && ok; ++b) {
            uint32_t *row = out + b * W;
            if (off[b] < 0 || off[b + 1] < off[b] || off[b + 1] > n_flat) {
                PyErr_SetString(PyExc_ValueError, "pack_csr: offsets must be non-decreasing and within the id array");
                ok = 0;
                break;
            }
            if (off[b + 1] == off[b]) { set_all(row, nE); continue; }
            for (int64_t i = off[b]; i < off[b + 1]; ++i)
                if (set_bit(row, (Py_ssize_t)flat[i], nE) < 0) { ok = 0; break; }
        }
    }
    PyBuffer_Release(&fb); PyBuffer_Release(&ofb); PyBuffer_Release(&ob);
    if (!ok) return NULL;
    Py_RETURN_NONE;
}

/* publish(flags, k, value): flags[k] = value with release ordering - everything this thread wrote
 * before (the chunk's masks) is visible to a thread that acquire-loads the flag, on any host ISA. */
static PyObject *publish(PyObject *self, PyObject *args) {
    PyObject *fobj;
    Py_ssize_t k;
    int value;
    if (!PyArg_ParseTuple(args, "Oni", &fobj, &k, &value)) return NULL;
    Py_buffer fb;
    if (PyObject_GetBuffer(fobj, &fb, PyBUF_WRITABLE | PyBUF_C_CONTIGUOUS) < 0) return NULL;
    if (fb.itemsize != 4 || k < 0 || k >= fb.len / 4) {
        PyBuffer_Release(&fb);
        PyErr_SetString(PyExc_ValueError, "publish: int32 flag array and an index inside it");
        return NULL;
    }
    __atomic_store_n((int32_t *)fb.buf + k, (int32_t)value, __ATOMIC_RELEASE);
    PyBuffer_Release(&fb);
    Py_RETURN_NONE;
}

static PyMethodDef methods[] = {
    {"publish", publish, METH_VARARGS, "publish(flags, k, value): release-store flags[k] = value (int32 array)"},
    {"pack_csr", pack_csr, METH_VARARGS, "pack_csr(flat_ids, offsets, nE, out): CSR id arrays -> uint32 subset bit masks"},
    {"pack", pack, METH_VARARGS, "pack(list_of_ids, nE, out): element-id sequences -> uint32 subset bit masks"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_cmpack", "subset bit-mask packer", -1, methods};

PyMODINIT_FUNC PyInit__cmpack(void) { return PyModule_Create(&moddef); }

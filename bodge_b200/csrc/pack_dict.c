/*
 * pack_dict.c -- host-side packing of the reference's dict API into the flat arrays bdg_scatter takes.
 *
 * The `with system as (H, D)` block hands the user two plain dicts keyed ((x,y,z),(x,y,z)) with 2x2 complex128
 * values (bodge/hamiltonian.py:86-89); the reference walks them entry by entry in Python in __exit__
 * (bodge/hamiltonian.py:102-118).  Here one C loop over PyDict_Next writes the six key integers and the four
 * complex values of every entry into caller-owned buffers, so that __exit__ is a copy instead of
 * `np.array(list(d.keys()))` (40 ms per 60 k entries).  Loaded with ctypes.PyDLL (the GIL is
 * held), separate from libbdg.so because it needs the CPython headers and no CUDA.
 *
 * Returns the number of entries written, or -(k + 1) when entry k is not of the plain form (key not a pair
 * of 3-tuples of integers, value not a C-contiguous 2x2 complex128 buffer): the caller then packs that dict
 * on the general numpy path.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <string.h>

static int key_ints(PyObject *coord, long long *out) {
    if (!PyTuple_Check(coord) || PyTuple_GET_SIZE(coord) != 3) return 0;
    for (int a = 0; a < 3; ++a) {
        PyObject *v = PyTuple_GET_ITEM(coord, a);
        if (PyFloat_Check(v)) return 0; /* no silent truncation of float coordinates */
        long long x = PyLong_AsLongLong(v);
        if (x == -1 && PyErr_Occurred()) {
            PyErr_Clear();
            return 0;
        }
        out[a] = x;
    }
    return 1;
}

static int value_copy(PyObject *val, double *out) {
    Py_buffer view;
    if (PyObject_GetBuffer(val, &view, PyBUF_C_CONTIGUOUS | PyBUF_FORMAT) != 0) {
        PyErr_Clear();
        return 0;
    }
    const int ok = view.len == 64 && view.itemsize == 16 && view.ndim == 2 && view.shape[0] == 2 && view.shape[1] == 2 &&
                   view.format && strcmp(view.format, "Zd") == 0;
    if (ok) memcpy(out, view.buf, 64);
    PyBuffer_Release(&view);
    return ok;
}

/* The same for a CubicLattice((Lx, Ly, Lz)): keys go straight to flat site indices z + Lz (y + Ly x)
 * (bodge/lattice.py:101-108).  An out-of-bounds coordinate also returns -(k + 1): the general path then raises the
 * reference's ValueError for it. */
long long bdg_pack_dict_cubic(PyObject *dict, long long Lx, long long Ly, long long Lz, int *site_i, int *site_j,
                              double *vals /* [n][2][2][2] */) {
    if (!PyDict_Check(dict)) return -1;
    Py_ssize_t pos = 0;
    PyObject *key, *val;
    long long n = 0, c[6];
    while (PyDict_Next(dict, &pos, &key, &val)) {
        if (!PyTuple_Check(key) || PyTuple_GET_SIZE(key) != 2) return -(n + 1);
        if (!key_ints(PyTuple_GET_ITEM(key, 0), c) || !key_ints(PyTuple_GET_ITEM(key, 1), c + 3)) return -(n + 1);
        for (int a = 0; a < 6; a += 3)
            if (c[a] < 0 || c[a] >= Lx || c[a + 1] < 0 || c[a + 1] >= Ly || c[a + 2] < 0 || c[a + 2] >= Lz) return -(n + 1);
        site_i[n] = (int)(c[2] + Lz * (c[1] + Ly * c[0]));
        site_j[n] = (int)(c[5] + Lz * (c[4] + Ly * c[3]));
        if (!value_copy(val, vals + 8 * n)) return -(n + 1);
        ++n;
    }
    return n;
}

long long bdg_pack_dict(PyObject *dict, long long *keys /* [n][2][3] */, double *vals /* [n][2][2][2] */) {
    if (!PyDict_Check(dict)) return -1;
    Py_ssize_t pos = 0;
    PyObject *key, *val;
    long long n = 0;
    while (PyDict_Next(dict, &pos, &key, &val)) {
        if (!PyTuple_Check(key) || PyTuple_GET_SIZE(key) != 2) return -(n + 1);
        if (!key_ints(PyTuple_GET_ITEM(key, 0), keys + 6 * n) || !key_ints(PyTuple_GET_ITEM(key, 1), keys + 6 * n + 3))
            return -(n + 1);
        if (!value_copy(val, vals + 8 * n)) return -(n + 1);
        ++n;
    }
    return n;
}

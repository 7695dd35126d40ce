/* Host-side helper of the QVM API mirror: the reference returns run()/run_and_measure() results as one Python
 * list of ints per trial (reference qvm_wavefunction.py:284-330).  For 10^6 trials x 33 qubits that is 3.3e7
 * list slots; numpy's tolist() builds them through its generic scalar path.  These two functions fill the row
 * lists directly (0 and 1 are the interpreter's cached small ints), straight from the sampled basis-state indices
 * where possible, which also skips the intermediate bit matrix.
 *
 * Loaded with ctypes.PyDLL (the GIL is held); built by build.py next to libqvm_b200.so.  Nothing here touches
 * the device. */
#include <Python.h>
#include <stdint.h>

/* rows[r][c] = bit sel[c] of idx[r] (sel[c] < 0 or > 63: always 0) */
PyObject* qvm_rows_from_indices(const uint64_t* idx, Py_ssize_t rows, const int32_t* sel, Py_ssize_t cols) {
    PyObject* bit[2] = {PyLong_FromLong(0), PyLong_FromLong(1)};
    PyObject* out = (bit[0] && bit[1]) ? PyList_New(rows) : NULL;
    for (Py_ssize_t r = 0; out && r < rows; ++r) {
        PyObject* row = PyList_New(cols);
        if (!row) {
            Py_CLEAR(out);
            break;
        }
        const uint64_t v = idx[r];
        for (Py_ssize_t c = 0; c < cols; ++c) {
            const int32_t s = sel[c];
            PyObject* o = bit[(s >= 0 && s < 64) ? (int)((v >> s) & 1u) : 0];
            Py_INCREF(o);
            PyList_SET_ITEM(row, c, o);
        }
        /* a list of ints cannot be part of a reference cycle: taken off the collector's books, a million of them cost
         * the cyclic garbage collector nothing (the list itself stays an ordinary list) */
        PyObject_GC_UnTrack(row);
        PyList_SET_ITEM(out, r, row);
    }
    Py_XDECREF(bit[0]);
    Py_XDECREF(bit[1]);
    return out;
}

/* rows[r][c] = m[r, c] of a strided int64 matrix (strides in bytes) */
PyObject* qvm_rows_from_int64(const char* m, Py_ssize_t rows, Py_ssize_t cols, Py_ssize_t row_stride,
                              Py_ssize_t col_stride) {
    PyObject* bit[2] = {PyLong_FromLong(0), PyLong_FromLong(1)};
    PyObject* out = (bit[0] && bit[1]) ? PyList_New(rows) : NULL;
    for (Py_ssize_t r = 0; out && r < rows; ++r) {
        PyObject* row = PyList_New(cols);
        if (!row) {
            Py_CLEAR(out);
            break;
        }
        PyList_SET_ITEM(out, r, row);
        for (Py_ssize_t c = 0; c < cols; ++c) {
            const int64_t v = *(const int64_t*)(m + r * row_stride + c * col_stride);
            PyObject* o;
            if (v == 0 || v == 1) {
                o = bit[v];
                Py_INCREF(o);
            } else {
                o = PyLong_FromLongLong(v);
                if (!o) {
                    for (Py_ssize_t k = c; k < cols; ++k) {        /* keep the row well-formed for the decref */
                        Py_INCREF(Py_None);
                        PyList_SET_ITEM(row, k, Py_None);
                    }
                    Py_CLEAR(out);
                    break;
                }
            }
            PyList_SET_ITEM(row, c, o);
        }
        if (out) PyObject_GC_UnTrack(row);
    }
    Py_XDECREF(bit[0]);
    Py_XDECREF(bit[1]);
    return out;
}

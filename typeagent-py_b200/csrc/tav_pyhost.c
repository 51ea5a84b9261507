/* Host-side helper for the Python mirror of VectorBase (not part of the C ABI in include/tavec.h).
 *
 * fuzzy_lookup_embedding_in_subset() receives the subset as a Python list of ints
 * (aitools/vectorbase.py:209-230).  Turning 1000 list items into an int64 buffer costs 18 us through
 * array('q', list) and 29 us through numpy — more than the GPU spends on the lookup — so this walks
 * the list once with the CPython API.  Loaded with ctypes.PyDLL (the GIL stays held); the Python
 * symbols resolve against the running interpreter, nothing here touches CUDA.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

/* Writes the items of `list` into out[0..cap).  Returns the item count, -1 when `list` is not an
 * exact list of exact ints that fit int64 (the caller falls back to the generic conversion, which
 * raises the reference's errors), -2 when it has more than `cap` items. */
long long tavhost_pack_int_list(PyObject* list, long long* out, long long cap)
{
    if (!PyList_CheckExact(list)) return -1;
    const Py_ssize_t n = PyList_GET_SIZE(list);
    if ((long long)n > cap) return -2;
    for (Py_ssize_t i = 0; i < n; ++i) {
        PyObject* item = PyList_GET_ITEM(list, i);
        if (!PyLong_CheckExact(item)) return -1;
#if PY_VERSION_HEX >= 0x030C0000
        if (PyUnstable_Long_IsCompact((const PyLongObject*)item)) {  /* one digit: no call, no overflow */
            out[i] = (long long)PyUnstable_Long_CompactValue((const PyLongObject*)item);
            continue;
        }
#endif
        int overflow = 0;
        const long long v = PyLong_AsLongLongAndOverflow(item, &overflow);
        if (overflow) return -1;
        out[i] = v;
    }
    return (long long)n;
}

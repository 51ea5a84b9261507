"""A CPU stand-in for libtavec's entry points that ``VectorBase`` calls at lookup time (test infrastructure).

It reads the query / subset / row-mask through the raw addresses the Python class passes and writes hits
through the output addresses, computing them with the oracle — so the HOST logic above the C ABI (argument
handling, paging, predicate pushdown, buffer reuse) is exercised by ``-m "not gpu"`` tests.  It is not a CPU
fallback of the product: it exists only here, and is installed by replacing ``VectorBase._ensure_device``."""

from __future__ import annotations

import ctypes as C

import numpy as np

from oracle import vectorbase_oracle as O
from typeagent_py_b200 import _capi


def _addr(p) -> int:
    if p is None:
        return 0
    if isinstance(p, int):
        return p
    return C.cast(p, C.c_void_p).value or 0


def _view(p, ctype, n):
    return np.ctypeslib.as_array(C.cast(_addr(p), C.POINTER(ctype)), (n,))


class FakeLib:
    def __init__(self, base):
        self.base = base            # the VectorBase whose host mirror holds the rows
        self.searches = []          # (n_queries, k, flags, subset_len) per tav_search call
        self.mask = None            # bool [N] from the last tav_set_row_mask
        self.mask_uploads = 0

    def tav_set_row_mask(self, ix, bits, n_rows, on_device, stream):
        words = _view(bits, C.c_uint32, (n_rows + 31) // 32).copy()
        self.mask = np.unpackbits(words.view(np.uint8), bitorder="little")[:n_rows].astype(bool)
        self.mask_uploads += 1
        return 0

    def tav_search(self, ix, qp, nq, k, floor, flags, sub_ptr, sub_len, item_offset, ip, sp, cp, stream):
        floor = float(getattr(floor, "value", floor))
        v = self.base._vectors
        dim = v.shape[1]
        q = _view(qp, C.c_float, nq * dim).reshape(nq, dim).copy()
        sub = _view(sub_ptr, C.c_int64, sub_len).copy() if _addr(sub_ptr) else None
        self.searches.append((nq, k, flags, None if sub is None else len(sub)))
        items = _view(ip, C.c_int64, nq * k).reshape(nq, k)
        scores = _view(sp, C.c_float, nq * k).reshape(nq, k)
        counts = _view(cp, C.c_int32, nq)
        for b in range(nq):
            if flags & _capi.TAV_USE_ROW_MASK:
                mask = self.mask
                hits = O.lookup(v, q[b], k, floor, predicate=lambda i: bool(mask[i]))
                if not flags & _capi.TAV_TIES_LOW_FIRST:      # the library's default order among equal scores
                    hits.sort(key=lambda h: (np.float32(h.score), h.item), reverse=True)
            elif sub is not None:
                hits = [O.Hit(int(sub[h.item]), h.score) for h in O.lookup(v[sub], q[b], k, floor)]
            else:
                hits = O.lookup(v, q[b], k, floor)
            counts[b] = len(hits)
            for j, h in enumerate(hits):
                items[b, j], scores[b, j] = h.item + item_offset, h.score
        return 0


def attach(base) -> FakeLib:
    """Route ``base``'s lookups to a FakeLib (its rows stay in the host mirror)."""
    fake = FakeLib(base)
    base._ensure_device = lambda: (fake, None)
    return fake

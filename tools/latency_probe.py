#!/usr/bin/env python3
"""Where does a single small lookup spend its time?  (host wall-clock, warm L2, 2000 iterations)"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import typeagent_py_b200 as tab  # noqa: E402
from typeagent_py_b200 import _capi  # noqa: E402


class Null:
    model_name = "probe"

    def add_embedding(self, k, e):
        pass


def bench(fn, n=2000, warm=200):
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n * 1e6


def main():
    lib = _capi.load()
    for rows, dim in ((1188, 1536), (10_000, 384), (100_000, 768)):
        rng = np.random.default_rng(0)
        v = rng.standard_normal((rows, dim)).astype(np.float32)
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        q = v[7].copy()
        base = tab.VectorBase(tab.TextEmbeddingIndexSettings(Null()))
        base.add_embeddings(None, v)
        base.fuzzy_lookup_embedding(q, 10, 0.0)
        ix = base._ix
        q2 = q.reshape(1, -1)
        items = np.empty((1, 10), np.int64)
        scores = np.empty((1, 10), np.float32)
        counts = np.empty(1, np.int32)
        qp, ip, sp, cp = (a.ctypes.data_as(C.c_void_p) for a in (q2, items, scores, counts))

        def raw():
            lib.tav_search(ix, qp, 1, 10, C.c_float(0.0), 0, None, 0, 0, ip, sp, cp, None)

        t_size = bench(lambda: lib.tav_size(ix))
        t_raw = bench(raw)
        t_arrays = bench(lambda: base.search_arrays(q2, 10, 0.0))
        t_lookup = bench(lambda: base.fuzzy_lookup_embedding(q, 10, 0.0))
        t_numpy = bench(lambda: np.dot(v, q), n=300, warm=30)
        sub = np.random.default_rng(99).choice(rows, size=min(1000, rows), replace=False).tolist()
        t_subset = bench(lambda: base.fuzzy_lookup_embedding_in_subset(q, sub, 10, 0.0))
        sub_np = np.asarray(sub, np.int64)
        subp = sub_np.ctypes.data_as(C.c_void_p)

        def raw_subset():
            lib.tav_search(ix, qp, 1, 10, C.c_float(0.0), 0, subp, len(sub_np), 0, ip, sp, cp, None)

        t_raw_subset = bench(raw_subset)
        t_aslist = bench(lambda: np.asarray(sub), n=300, warm=30)
        base.enable_timing()            # GPU-side duration of the one kernel (events add host time: separate pass)
        gpu_ms = []
        for _ in range(200):
            base.fuzzy_lookup_embedding(q, 10, 0.0)
            gpu_ms.append(base.last_timing()["scan_ms"])
        launches = base.last_timing()["launches"]
        base.enable_timing(False)
        gpu_us = sorted(gpu_ms)[len(gpu_ms) // 2] * 1e3
        print(f"{rows}x{dim}: ctypes call {t_size:.1f} us | tav_search (raw ctypes) {t_raw:.1f} us | "
              f"search_arrays {t_arrays:.1f} us | fuzzy_lookup_embedding {t_lookup:.1f} us | in_subset(1000) {t_subset:.1f} us (raw ctypes {t_raw_subset:.1f} us) "
              f"(np.asarray(list) alone {t_aslist:.1f} us) | kernel on the GPU {gpu_us:.1f} us ({launches} "
              f"launch) | np.dot alone {t_numpy:.1f} us", flush=True)


if __name__ == "__main__":
    main()

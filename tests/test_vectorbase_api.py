"""Host-side behaviour of the GPU VectorBase / EmbeddingIndex that needs no device:
the semantics of the reference's tests/test_vectorbase.py (add / serialize / clear / errors /
settings) run against BOTH the oracle restatement and the product class, plus the C-ABI
surface check (library loads, every symbol of include/tavec.h is exported, and compute entry
points fail loudly without a GPU).  Lookups are in tests/test_gpu_parity.py (``-m gpu``).
"""

from __future__ import annotations

import asyncio
import ctypes as C
import os
import re
from types import SimpleNamespace

import numpy as np
import pytest

import typeagent_py_b200 as tab
from oracle import vectorbase_oracle as O
from typeagent_py_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_oracle():
    return O.OracleVectorBase(SimpleNamespace(embedding_model=O.FakeEmbeddingModel(),
                                              min_score=0.85, max_matches=None))


def make_gpu():
    return tab.VectorBase(tab.TextEmbeddingIndexSettings(O.FakeEmbeddingModel()))


BACKENDS = {"oracle": make_oracle, "gpu_class": make_gpu}


@pytest.fixture(params=list(BACKENDS))
def make_base(request):
    return BACKENDS[request.param]


SAMPLES = {
    "word1": np.array([0.1, 0.2, 0.3], dtype=np.float32),
    "word2": np.array([0.4, 0.5, 0.6], dtype=np.float32),
    "word3": np.array([0.7, 0.8, 0.9], dtype=np.float32),
}


def test_add_embedding(make_base):
    base = make_base()
    for key, e in SAMPLES.items():
        base.add_embedding(key, e)
    assert len(base) == 3
    for i, e in enumerate(SAMPLES.values()):
        np.testing.assert_array_equal(base.serialize_embedding_at(i), e)


def test_add_embeddings_matches_single_adds_and_fills_cache(make_base):
    one, bulk = make_base(), make_base()
    keys = list(SAMPLES)
    for key, e in SAMPLES.items():
        one.add_embedding(key, e)
    bulk.add_embeddings(keys, np.stack([SAMPLES[k] for k in keys]))
    assert len(one) == len(bulk)
    np.testing.assert_array_equal(one.serialize(), bulk.serialize())
    assert set(one._model._cache) == set(bulk._model._cache) == set(keys)
    for k in keys:
        np.testing.assert_array_equal(one._model._cache[k], bulk._model._cache[k])


def test_add_key_and_keys_with_and_without_cache(make_base):
    base = make_base()
    for key in SAMPLES:
        asyncio.run(base.add_key(key))
    assert len(base) == 3 and set(base._model._cache) == set(SAMPLES)
    nocache = make_base()
    for key in SAMPLES:
        asyncio.run(nocache.add_key(key, cache=False))
    assert len(nocache) == 3 and nocache._model._cache == {}
    many = make_base()
    got = asyncio.run(many.add_keys(list(SAMPLES), cache=False))
    assert got.shape == (3, 3) and len(many) == 3 and many._model._cache == {}
    assert asyncio.run(many.add_keys([])) is None


def test_clear_serialize_deserialize_roundtrip(make_base):
    base = make_base()
    for key, e in SAMPLES.items():
        base.add_embedding(key, e)
    blob = base.serialize()
    assert blob.dtype == np.float32 and blob.shape == (3, 3)
    other = make_base()
    other.deserialize(blob)
    assert len(other) == 3
    for i in range(3):
        np.testing.assert_array_equal(other.serialize_embedding_at(i), base.serialize_embedding_at(i))
    other.deserialize(None)
    assert len(other) == 0
    base.clear()
    assert len(base) == 0 and base.serialize().shape == (0, 3)
    empty = make_base()
    empty.deserialize(np.zeros((0, 7), np.float32))  # cannot fix the width: just clears
    assert len(empty) == 0 and empty._embedding_size == 0
    assert empty.serialize().shape == (0,)


def test_deserialize_adopts_without_copy_and_append_does_not_write_into_it(make_base):
    base = make_base()
    data = np.arange(12, dtype=np.float32).reshape(4, 3)
    base.deserialize(data)
    assert base.serialize() is data or np.shares_memory(base.serialize(), data)
    base.add_embedding(None, [9, 9, 9])
    assert len(base) == 5 and data.shape == (4, 3)
    np.testing.assert_array_equal(data, np.arange(12, dtype=np.float32).reshape(4, 3))
    np.testing.assert_array_equal(base.serialize()[4], [9, 9, 9])


def test_bool_len_get_embedding_at(make_base):
    base = make_base()
    assert bool(base) is True and len(base) == 0
    for key, e in SAMPLES.items():
        base.add_embedding(key, e)
    for i, e in enumerate(SAMPLES.values()):
        np.testing.assert_array_equal(base.get_embedding_at(i), e)
    with pytest.raises(IndexError):
        base.get_embedding_at(3)
    with pytest.raises(IndexError):
        base.get_embedding_at(-1)
    assert base.serialize_embedding_at(3) is None


def test_size_and_ndim_errors(make_base):
    base = make_base()
    base.add_embedding(None, np.array([0.1, 0.2, 0.3], np.float32))
    with pytest.raises(ValueError, match="Embedding size mismatch"):
        base.add_embedding(None, np.zeros(5, np.float32))
    with pytest.raises(ValueError, match="Embedding size mismatch"):
        base.add_embeddings(None, np.zeros((1, 5), np.float32))
    with pytest.raises(ValueError, match="Expected 2D"):
        base.add_embeddings(None, np.zeros(3, np.float32))


def test_many_appends_are_amortised_and_exact():
    base = make_gpu()
    rng = np.random.default_rng(0)
    rows = rng.standard_normal((1000, 16)).astype(np.float32)
    for r in rows[:500]:
        base.add_embedding(None, r)
    base.add_embeddings(None, rows[500:])
    np.testing.assert_array_equal(base.serialize(), rows)
    snapshot = base.serialize()
    base.add_embedding(None, rows[0])
    assert snapshot.shape == (1000, 16) and len(base) == 1001  # old view is a stable snapshot


@pytest.mark.parametrize(
    ("model_name", "expected"),
    [("text-embedding-3-large", 0.74), ("text-embedding-3-small", 0.73),
     ("text-embedding-ada-002", 0.93), ("custom-embedding-model", 0.85)],
)
def test_settings_defaults(model_name, expected):
    model = SimpleNamespace(model_name=model_name)
    s = tab.TextEmbeddingIndexSettings(embedding_model=model)
    assert s.min_score == expected and s.max_matches is None and s.batch_size == 8
    s = tab.TextEmbeddingIndexSettings(embedding_model=model, min_score=0.55, max_matches=7, batch_size=3)
    assert (s.min_score, s.max_matches, s.batch_size) == (0.55, 7, 3)
    assert tab.TextEmbeddingIndexSettings(embedding_model=model, max_matches=0).max_matches is None
    assert tab.DEFAULT_MIN_SCORE == 0.85


def test_embedding_index_wrapper_host_side():
    idx = tab.EmbeddingIndex(tab.TextEmbeddingIndexSettings(O.FakeEmbeddingModel()),
                             np.eye(3, dtype=np.float32))
    assert len(idx) == 3 and asyncio.run(idx.size()) == 3 and not asyncio.run(idx.is_empty())
    idx.push(np.ones((2, 3), np.float32))
    np.testing.assert_array_equal(idx.get(4), [1, 1, 1])
    blob = idx.serialize()
    other = tab.EmbeddingIndex(tab.TextEmbeddingIndexSettings(O.FakeEmbeddingModel()))
    other.deserialize(blob)
    assert len(other) == 5
    with pytest.raises(AssertionError):
        other.deserialize(blob.astype(np.float64))
    with pytest.raises(AssertionError):
        other.deserialize(np.zeros((2, 4), np.float32))
    idx.clear()
    assert len(idx) == 0


# ------------------------------------------------------------------ the C ABI surface
def declared_symbols() -> list[str]:
    with open(os.path.join(ROOT, "include", "tavec.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(tav_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _capi.load()
    names = declared_symbols()
    assert len(names) >= 17, names
    for name in names:
        assert hasattr(lib, name), f"libtavec.so does not export {name}"
        assert name in _capi.SIGNATURES, f"ctypes binding lacks {name}"
    assert set(_capi.SIGNATURES) == set(names)
    assert lib.tav_abi_version() == _capi.ABI_VERSION == 2


def test_no_gpu_means_loud_failure_not_cpu_fallback():
    if _capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    lib = _capi.load()
    handle = C.c_void_p()
    rc = lib.tav_create(0, 8, _capi.TAV_F32, 0, 0, C.byref(handle))
    assert rc == _capi.TAV_ERR_CUDA and "no CPU fallback" in _capi.last_error()
    base = make_gpu()
    base.add_embedding(None, [1.0, 0.0])
    with pytest.raises(RuntimeError, match="no CUDA device"):
        base.fuzzy_lookup_embedding(np.array([1.0, 0.0], np.float32))
    with pytest.raises(RuntimeError):
        base.fuzzy_lookup_embedding_in_subset(np.array([1.0, 0.0], np.float32), [0])


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "typeagent-py_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, name)) as f:
                    text = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), name


def test_host_helper_packs_int_lists_and_declines_everything_else():
    """libtavhost.so (csrc/tav_pyhost.c) only short-cuts `list[int] -> int64 buffer` for
    fuzzy_lookup_embedding_in_subset; anything else must be declined (-1) so that the generic numpy
    conversion — which raises the reference's errors — handles it."""
    pack = _capi.pack_int_list()
    if pack is None:
        pytest.skip("libtavhost.so not built (no Python.h / gcc)")
    buf = np.full(8, -7, np.int64)
    assert pack([3, 0, -5, 2**62], buf.ctypes.data, 8) == 4
    assert buf[:4].tolist() == [3, 0, -5, 2**62] and buf[4] == -7
    assert pack([], buf.ctypes.data, 8) == 0
    for declined in ([1, True], (1, 2), [1, 2**70], [1.0], [np.int64(3)], "ab", None):
        assert pack(declined, buf.ctypes.data, 8) == -1
    assert pack(list(range(9)), buf.ctypes.data, 8) == -2


def test_single_lookup_host_path_against_a_fake_library():
    """Host logic of the latency path (`_lookup_one`) without a GPU: a stand-in for libtavec.tav_search reads
    the query / subset through the raw addresses the class passes and writes hits through the output
    addresses — pointer extraction, the CPython-API subset packing and its fallbacks, buffer reuse, k clamp."""
    rng = np.random.default_rng(5)
    v = rng.standard_normal((50, 8)).astype(np.float32)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    base = make_gpu()
    base.add_embeddings(None, v)
    seen = []

    class FakeLib:
        @staticmethod
        def tav_search(ix, qp, nq, k, floor, flags, sub_ptr, sub_len, item_offset, ip, sp, cp, stream):
            q = np.ctypeslib.as_array(C.cast(qp, C.POINTER(C.c_float)), (8,)).copy()
            sub = None if not sub_ptr else np.ctypeslib.as_array(C.cast(sub_ptr, C.POINTER(C.c_int64)), (sub_len,)).copy()
            seen.append((q, sub, k))
            rows = v if sub is None else v[sub]
            want = O.lookup(rows, q, k, float(floor))
            items = np.ctypeslib.as_array(C.cast(ip, C.POINTER(C.c_int64)), (k,))
            scores = np.ctypeslib.as_array(C.cast(sp, C.POINTER(C.c_float)), (k,))
            for j, h in enumerate(want):
                items[j], scores[j] = (h.item if sub is None else int(sub[h.item])), h.score
            C.cast(cp, C.POINTER(C.c_int32))[0] = len(want)
            return 0

    base._ensure_device = lambda: (FakeLib, None)
    q = v[3].copy()
    got = base.fuzzy_lookup_embedding(q, 4, 0.0)
    assert [h.item for h in got] == [h.item for h in O.lookup(v, q, 4, 0.0)] and got[0].item == 3
    ro = q.copy()
    ro.flags.writeable = False                      # read-only query: the slower pointer path
    assert [h.item for h in base.fuzzy_lookup_embedding(ro, 4, 0.0)] == [h.item for h in got]
    assert base.fuzzy_lookup_embedding(list(map(float, q)), 2, 0.0)[0].item == 3   # any float sequence
    # subsets: exact list of ints (packed by libtavhost when built), numpy ints and bools (generic path), arrays
    for subset in ([7, 3, 11, 3], [np.int64(7), 3, 11], [7, True, 3], np.array([7, 3, 11], np.int32)):
        hits = base.fuzzy_lookup_embedding_in_subset(q, subset, 10, 0.0)
        expect = [int(x) for x in np.asarray(subset, dtype=np.int64)]
        assert seen[-1][1].tolist() == expect and seen[-1][2] == len(expect)      # k clamped to the subset
        assert hits[0].item == 3 and {h.item for h in hits} <= set(expect)
    with pytest.raises(IndexError):
        base.fuzzy_lookup_embedding_in_subset(q, [1.5, 2.0], 3, 0.0)
    with pytest.raises(ValueError):
        base.fuzzy_lookup_embedding(np.zeros(9, np.float32), 3, 0.0)
    assert base.fuzzy_lookup_embedding(q, 3, float("nan")) == []
    assert base.fuzzy_lookup_embedding_in_subset(q, [], 3, 0.0) == []
    # a longer list than the reusable buffer grows it
    big = list(range(50)) * 100
    assert len(base.fuzzy_lookup_embedding_in_subset(q, big, 5, 0.0)) == 5 and seen[-1][1].tolist() == big

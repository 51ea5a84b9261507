"""Randomised parity (hypothesis): arbitrary small shapes, k, thresholds, subsets and storage
dtypes against the CPU oracle — the ragged / degenerate corners seeded tests do not enumerate."""

from __future__ import annotations

import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import typeagent_py_b200 as tab
from oracle import vectorbase_oracle as O
from tests.parity import assert_hits_match

pytestmark = pytest.mark.gpu

COMMON = dict(deadline=None, max_examples=60, suppress_health_check=[HealthCheck.too_slow])


def base_for(v, storage="float32", path=None):
    b = tab.VectorBase(tab.TextEmbeddingIndexSettings(O.FakeEmbeddingModel()), storage_dtype=storage)
    if len(v):
        b.add_embeddings(None, v)
    b.force_path = path
    return b


@settings(**COMMON)
@given(n=st.integers(0, 2500), d=st.integers(1, 260), seed=st.integers(0, 2**31 - 1),
       k=st.one_of(st.none(), st.integers(0, 70)), ms=st.one_of(st.none(), st.floats(0.0, 0.7)),
       nq=st.integers(1, 11))
def test_fp32_lookup(n, d, seed, k, ms, nq):
    v, q = O.make_corpus(max(n, 1), d, seed, n_queries=nq)
    v = v[:n]
    base = base_for(v)
    got = base.fuzzy_lookup_embeddings(q, max_hits=k, min_score=ms)
    for qq, hits in zip(q, got):
        assert_hits_match(hits, O.lookup(v, qq, k, ms), min_score=ms, what=f"n={n} d={d} k={k} ms={ms}")


@settings(**COMMON)
@given(n=st.integers(1, 1500), d=st.integers(1, 130), seed=st.integers(0, 2**31 - 1), k=st.integers(1, 40),
       ms=st.floats(0.0, 0.6), data=st.data())
def test_subset_lookup(n, d, seed, k, ms, data):
    v, q = O.make_corpus(n, d, seed)
    subset = data.draw(st.lists(st.integers(-n, n - 1), min_size=0, max_size=200))
    base = base_for(v)
    got = base.fuzzy_lookup_embedding_in_subset(q[0], subset, k, ms)
    want = O.lookup_in_subset(v, q[0], subset, k, ms)
    # the reference returns the caller's ordinal (possibly negative); duplicates allowed
    assert_hits_match(got, want, min_score=ms)


@settings(**dict(COMMON, max_examples=40))
@given(n=st.integers(1, 4000), d8=st.integers(1, 40), seed=st.integers(0, 2**31 - 1), k=st.integers(1, 50),
       nq=st.integers(1, 300), storage=st.sampled_from(["bfloat16", "float16"]),
       path=st.sampled_from(["scan", "mma"]), ms=st.sampled_from([0.0, 0.5, 0.55]))
def test_16bit_storage_both_paths(n, d8, seed, k, nq, storage, path, ms):
    d = d8 * 8  # the tensor-core path needs 16-byte rows
    v, q = O.make_corpus(n, d, seed, n_queries=nq)
    vr, qr = O.round_to_storage(v, storage), O.round_to_storage(q, storage)
    base = base_for(v, storage, path)
    got = base.fuzzy_lookup_embeddings(qr, max_hits=k, min_score=ms)
    for i in {0, nq // 2, nq - 1}:
        assert_hits_match(got[i], O.lookup(vr, qr[i], k, ms), min_score=ms, what=f"{storage}/{path} n={n} d={d} k={k}")

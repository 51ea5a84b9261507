"""Host logic of the Python ``VectorBase`` above the C ABI, on CPU: the lookups are routed to
``tests/fake_lib.py`` (oracle arithmetic behind libtavec's call signatures), so what is tested here is
everything the class itself decides — defaults and quirks of ``fuzzy_lookup_embedding`` (reference
aitools/vectorbase.py:163-201), the predicate pushdown (one masked search, cached bitmask, the page-first
strategy of large indexes), subset handling (:203-230) and the batched entry points."""

from __future__ import annotations

import asyncio

import numpy as np
import pytest

import typeagent_py_b200 as tab
from oracle import vectorbase_oracle as O
from tests.fake_lib import attach
from tests.parity import assert_hits_match
from typeagent_py_b200 import _capi


def make(v, **settings):
    base = tab.VectorBase(tab.TextEmbeddingIndexSettings(O.FakeEmbeddingModel(), **settings))
    base.add_embeddings(None, v)
    return base, attach(base)


def test_defaults_and_the_max_hits_quirks():
    v, q = O.make_corpus(300, 16, seed=1, n_queries=2)
    base, fake = make(v)
    assert len(base.fuzzy_lookup_embedding(q[0])) == 10                      # max_hits None -> 10 (:170-171)
    assert fake.searches[-1][1] == 10
    everything = base.fuzzy_lookup_embedding(q[0], 0, 0.55)                   # 0 -> every passing row (quirk Q2)
    assert_hits_match(everything, O.lookup(v, q[0], len(v), 0.55), min_score=0.55)
    assert fake.searches[-1][1] == len(v)
    assert base.fuzzy_lookup_embedding(q[0], 0, 0.0, predicate=lambda i: True) == []   # predicate path: [:0]
    with pytest.raises(ValueError):
        base.fuzzy_lookup_embedding(q[0], -1)
    assert len(base.fuzzy_lookup_embedding(q[0], 1000, 0.0)) == len(v)      # k clamped to the rows
    assert fake.searches[-1][1] == len(v)


def test_predicate_is_one_masked_search_and_the_mask_is_cached():
    v, q = O.make_corpus(500, 12, seed=2, n_queries=3)
    base, fake = make(v)
    calls = []

    def pred(i):
        calls.append(i)
        return i % 3 == 0

    for qi in range(3):
        got = base.fuzzy_lookup_embedding(q[qi], 7, 0.3, predicate=pred)
        want = O.lookup(v, q[qi], 7, 0.3, predicate=lambda i: i % 3 == 0)
        assert [h.item for h in got] == [h.item for h in want]
    assert len(calls) == len(v)                              # evaluated once per row, not once per lookup
    assert fake.mask_uploads == 1 and len(fake.searches) == 3
    assert all(f & _capi.TAV_USE_ROW_MASK and f & _capi.TAV_TIES_LOW_FIRST for _, _, f, _ in fake.searches)
    # appended rows invalidate the cached mask (the key carries the row count)
    base.add_embedding(None, q[0])
    got = base.fuzzy_lookup_embedding(q[0], 3, 0.0, predicate=pred)
    assert len(calls) == 2 * len(v) + 1 and fake.mask_uploads == 2
    assert got[0].item == (500 if 500 % 3 == 0 else got[0].item)
    # a different predicate object -> its own mask; clear_predicate_cache() forgets them all
    base.fuzzy_lookup_embedding(q[1], 3, 0.0, predicate=lambda i: i < 10)
    assert fake.mask_uploads == 3
    base.clear_predicate_cache()
    base.fuzzy_lookup_embedding(q[1], 3, 0.0, predicate=pred)
    assert fake.mask_uploads == 4


def test_predicate_ties_come_back_lower_ordinal_first():
    row = O.make_corpus(1, 8, seed=3)[0]
    v = np.repeat(row, 40, axis=0)
    base, fake = make(v)
    got = base.fuzzy_lookup_embedding(row[0], 5, 0.0, predicate=lambda i: i % 4 == 2)
    assert [h.item for h in got] == [2, 6, 10, 14, 18]      # the reference's stable sort (:199-200)


def test_large_index_tries_one_unfiltered_page_before_building_the_mask(monkeypatch):
    v, q = O.make_corpus(900, 8, seed=4, n_queries=2)
    base, fake = make(v)
    monkeypatch.setattr(tab.VectorBase, "_PREDICATE_MASK_ROWS", 256)   # "large" starts here for this test
    loose = lambda i: i % 2 == 0  # noqa: E731
    got = base.fuzzy_lookup_embedding(q[0], 5, 0.0, predicate=loose)
    assert [h.item for h in got] == [h.item for h in O.lookup(v, q[0], 5, 0.0, predicate=loose)]
    assert len(fake.searches) == 1 and not fake.searches[0][2] & _capi.TAV_USE_ROW_MASK and fake.mask_uploads == 0
    tight = lambda i: i in (17, 400, 899)  # noqa: E731  (the page cannot settle it)
    got = base.fuzzy_lookup_embedding(q[1], 5, 0.0, predicate=tight)
    assert [h.item for h in got] == [h.item for h in O.lookup(v, q[1], 5, 0.0, predicate=tight)]
    assert len(fake.searches) == 3 and fake.searches[2][2] & _capi.TAV_USE_ROW_MASK and fake.mask_uploads == 1
    base.fuzzy_lookup_embedding(q[0], 5, 0.0, predicate=tight)          # mask cached: straight to it
    assert len(fake.searches) == 4 and fake.mask_uploads == 1


def test_subset_lookups_map_back_to_ordinals_and_validate_on_the_host():
    v, q = O.make_corpus(200, 10, seed=5, n_queries=1)
    base, fake = make(v)
    subset = [5, 5, 199, -1, 0, -200, 77]
    got = base.fuzzy_lookup_embedding_in_subset(q[0], subset, 4, 0.0)
    want = O.lookup_in_subset(v, q[0], subset, 4, 0.0)
    assert_hits_match(got, want)
    assert fake.searches[-1][3] == len(subset)
    items, scores, counts = base.search_arrays(q, 3, 0.0, subset=np.array([9, 8, 7], np.int32))
    assert set(items[0, : counts[0]].tolist()) <= {7, 8, 9} and counts[0] == 3
    with pytest.raises(IndexError):
        base.search_arrays(q, 3, 0.0, subset=[0.5, 1.5])
    with pytest.raises(ValueError):
        base.search_arrays(q, 3, 0.0, subset=[1], allowed=np.ones(200, bool))
    with pytest.raises(ValueError):
        base.search_arrays(q, 3, 0.0, allowed=np.ones(199, bool))      # wrong mask length
    with pytest.raises(ValueError):
        base.search_arrays(q, 0, 0.0)


def test_batched_entry_points_and_preallocated_outputs():
    v, q = O.make_corpus(400, 3, seed=6, n_queries=5)      # FakeEmbeddingModel embeds keys in 3 dimensions
    base, fake = make(v, min_score=0.4, max_matches=3)
    lists = base.fuzzy_lookup_embeddings(q, 4, 0.5)
    assert len(fake.searches) == 1 and fake.searches[0][0] == 5          # ONE search for the batch
    for b in range(5):
        assert_hits_match(lists[b], O.lookup(v, q[b], 4, 0.5), min_score=0.5)
    out = (np.empty((5, 4), np.int64), np.empty((5, 4), np.float32), np.empty(5, np.int32))
    items, scores, counts = base.search_arrays(q, 4, 0.5, out=out)
    assert items is out[0] and counts is out[2]
    with pytest.raises(ValueError):
        base.search_arrays(q, 4, 0.5, out=(np.empty((5, 3), np.int64), out[1], out[2]))
    assert base.search_arrays(q, 4, float("nan"))[2].tolist() == [0] * 5 and len(fake.searches) == 2
    # fuzzy_lookup: defaults from the settings (max_matches, min_score), embedding from the model (:232-246)
    hits = asyncio.run(base.fuzzy_lookup("some key"))
    assert fake.searches[-1][1] == 3 and all(h.score >= np.float32(0.4) for h in hits)
    many = asyncio.run(base.fuzzy_lookup_keys(["a", "b", "c"], max_hits=2))
    assert len(many) == 3 and fake.searches[-1][0] == 3 and fake.searches[-1][1] == 2

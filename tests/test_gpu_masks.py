"""Predicate / post-filter pushdown, the single-launch latency form of the row
scan, and the device-side chunk -> message fold — against the oracle and, where they are vendored
(``oracle/_ref``), the reference's own classes.

Reference semantics: ``VectorBase.fuzzy_lookup_embedding(predicate=...)`` (aitools/vectorbase.py:
191-201): rows >= min_score that pass the predicate, stable sort by descending score, first k;
``MessageTextIndex.to_scored_message_ordinals`` (storage/memory/messageindex.py:185-207).
"""

from __future__ import annotations

import numpy as np
import pytest

import typeagent_py_b200 as tab
from oracle import vectorbase_oracle as O
from tests.parity import assert_hits_match

pytestmark = pytest.mark.gpu


def make_base(v, storage="float32", path=None):
    base = tab.VectorBase(tab.TextEmbeddingIndexSettings(O.FakeEmbeddingModel()), storage_dtype=storage)
    base.add_embeddings(None, v)
    base.force_path = path
    return base


def count_searches(base):
    calls = []
    inner = base.search_arrays

    def wrapped(*a, **k):
        calls.append(k)
        return inner(*a, **k)

    base.search_arrays = wrapped
    return calls


@pytest.mark.parametrize("n,k,ms,mod", [(5000, 10, 0.0, 3), (5000, 50, 0.5, 7), (20000, 10, 0.0, 1000),
                                        (1188, 5, 0.45, 2), (300, 400, 0.0, 2)])
def test_predicate_is_one_masked_search_with_reference_order(n, k, ms, mod):
    v, q = O.make_corpus(n, 64, seed=n + k, n_queries=3)
    base = make_base(v)
    calls = count_searches(base)
    pred = lambda i: i % mod == 0  # noqa: E731
    for qq in q:
        got = base.fuzzy_lookup_embedding(qq, k, ms, predicate=pred)
        want = O.lookup(v, qq, k, ms, predicate=pred)
        assert_hits_match(got, want, min_score=ms, what=f"predicate n={n} mod={mod}")
        assert all(h.item % mod == 0 for h in got)
    assert len(calls) == 3 and all("allowed" in c for c in calls)   # ONE search per lookup, mask cached


def test_predicate_ties_keep_the_reference_order_lower_ordinal_first():
    """Exactly equal scores: the reference's stable sort keeps ascending ordinals (vectorbase.py:200),
    also across the rank-k boundary; the masked scan reproduces it bit for bit."""
    row = O.make_corpus(1, 32, seed=2)[0]
    v = np.repeat(row, 4000, axis=0)
    base = make_base(v)
    got = base.fuzzy_lookup_embedding(row[0], 6, 0.0, predicate=lambda i: i % 5 == 1)
    want = O.lookup(v, row[0], 6, 0.0, predicate=lambda i: i % 5 == 1)
    assert [h.item for h in got] == [h.item for h in want] == [1, 6, 11, 16, 21, 26]


def test_predicate_max_hits_zero_and_empty_results():
    v, q = O.make_corpus(2000, 48, seed=8, n_queries=1)
    base = make_base(v)
    assert base.fuzzy_lookup_embedding(q[0], 0, 0.0, predicate=lambda i: True) == []   # reference: [:0]
    assert O.lookup(v, q[0], 0, 0.0, predicate=lambda i: True) == []
    assert base.fuzzy_lookup_embedding(q[0], 5, 0.0, predicate=lambda i: False) == []
    assert base.fuzzy_lookup_embedding(q[0], 5, 0.999, predicate=lambda i: True) == []


def test_predicate_cache_survives_recycled_ids_and_can_be_cleared():
    v, q = O.make_corpus(3000, 32, seed=12, n_queries=1)
    base = make_base(v)
    # many short-lived predicates (the cache holds 8; CPython hands a collected lambda's id to the next one)
    for m in range(2, 24):
        got = base.fuzzy_lookup_embedding(q[0], 7, 0.0, predicate=lambda i, m=m: i % m == 1)
        want = O.lookup(v, q[0], 7, 0.0, predicate=lambda i, m=m: i % m == 1)
        assert [h.item for h in got] == [h.item for h in want], m
    # a predicate whose meaning changes between lookups: cached per function object until told otherwise
    allowed = {5, 17, 300}
    pred = lambda i: i in allowed  # noqa: E731
    assert sorted(h.item for h in base.fuzzy_lookup_embedding(q[0], 10, 0.0, predicate=pred)) == [5, 17, 300]
    allowed.add(1234)
    base.clear_predicate_cache()
    assert sorted(h.item for h in base.fuzzy_lookup_embedding(q[0], 10, 0.0, predicate=pred)) == [5, 17, 300, 1234]


def test_large_index_predicate_tries_one_page_then_the_mask():
    v, q = O.make_corpus(70000, 32, seed=5, n_queries=2)
    base = make_base(v)
    calls = count_searches(base)
    loose = lambda i: i % 2 == 0  # noqa: E731  (settles on the first page)
    assert_hits_match(base.fuzzy_lookup_embedding(q[0], 10, 0.0, predicate=loose),
                      O.lookup(v, q[0], 10, 0.0, predicate=loose))
    assert len(calls) == 1 and "allowed" not in calls[0]
    tight = lambda i: i % 9973 == 5  # noqa: E731  (7 rows pass: the page cannot settle it)
    assert_hits_match(base.fuzzy_lookup_embedding(q[1], 10, 0.0, predicate=tight),
                      O.lookup(v, q[1], 10, 0.0, predicate=tight))
    assert len(calls) == 3 and "allowed" in calls[2]
    base.fuzzy_lookup_embedding(q[0], 10, 0.0, predicate=tight)       # mask cached: straight to it
    assert len(calls) == 4 and "allowed" in calls[3]


@pytest.mark.parametrize("storage,path", [("float32", "scan"), ("float32", "scan2"), ("bfloat16", "mma"),
                                          ("float32", "mma")])
def test_row_mask_on_every_kernel_path(storage, path):
    n, d, b, k = 40000, 128, 40, 20
    v, q = O.make_corpus(n, d, seed=11, n_queries=b)
    if storage != "float32":
        v, q = O.round_to_storage(v, storage), O.round_to_storage(q, storage)
    rng = np.random.default_rng(3)
    allowed = rng.random(n) < 0.3
    base = make_base(v, storage, path)
    items, scores, counts = base.search_arrays(q, k, 0.0, allowed=allowed)
    pred = lambda i: bool(allowed[i])  # noqa: E731
    for i in range(0, b, 5):
        got = {"items": items[i, : counts[i]].tolist(), "scores": scores[i, : counts[i]].tolist()}
        assert all(allowed[r] for r in got["items"])
        assert_hits_match(got, O.lookup(v, q[i], k, 0.0, predicate=pred), what=f"{storage}/{path} q{i}")
    # a very selective mask starves the sampled threshold: the exact fallback must honour the mask too
    few = np.zeros(n, bool)
    few[[5, 77, 30001, 39999]] = True
    items, scores, counts = base.search_arrays(q[:17], k, 0.0, allowed=few)
    for i in range(17):
        assert sorted(items[i, : counts[i]].tolist()) == [5, 77, 30001, 39999]
    # and no mask again afterwards
    items, scores, counts = base.search_arrays(q[:17], k, 0.0)
    assert_hits_match({"items": items[3, : counts[3]].tolist(), "scores": scores[3, : counts[3]].tolist()},
                      O.lookup(v, q[3], k, 0.0))


def test_single_launch_form_matches_the_two_kernel_form_and_the_oracle():
    """One host query -> ONE kernel launch (query in the kernel parameters, last CTA merges)."""
    for n, d, k in [(1000, 384, 10), (10000, 384, 10), (1188, 1536, 50), (70000, 64, 10), (37, 8, 10), (5000, 100, 3)]:
        v, q = O.make_corpus(n, d, seed=n + d, n_queries=4)
        one = make_base(v, "float32", None)
        two = make_base(v, "float32", "scan2")
        for qq in q:
            a = one.fuzzy_lookup_embedding(qq, k, 0.0)
            assert one.last_timing()["launches"] == 1 and one.last_timing()["path"] == "scan"
            b = two.fuzzy_lookup_embedding(qq, k, 0.0)
            assert two.last_timing()["launches"] == 2
            assert [(h.item, h.score) for h in a] == [(h.item, h.score) for h in b]
            assert_hits_match(a, O.lookup(v, qq, k, 0.0), what=f"n={n} d={d}")
        sub = list(range(0, n, 3)) + [0, -1]
        assert_hits_match(one.fuzzy_lookup_embedding_in_subset(q[0], sub, k, 0.0),
                          O.lookup_in_subset(v, q[0], sub, k, 0.0), what="subset in the kernel parameters")
        assert one.last_timing()["launches"] == 1   # short subsets ride in the parameters; longer ones are uploaded,
        #                                               but an L2-sized scan still ends in the scan kernel itself
        with pytest.raises(IndexError):
            one.fuzzy_lookup_embedding_in_subset(q[0], [0, n], k, 0.0)


def test_device_fold_of_chunk_hits_to_messages():
    import torch

    from typeagent_py_b200.formats import fold_chunk_hits_to_messages

    n, d, b, k = 30000, 64, 33, 40
    v, q = O.make_corpus(n, d, seed=21, n_queries=b)
    vr, qr = O.round_to_bfloat16(v), O.round_to_bfloat16(q)
    chunk_to_message = (np.arange(n) // 7).astype(np.int32)         # 7 chunks per message
    base = make_base(v, "bfloat16", "mma")
    base._ensure_device()
    qd = torch.from_numpy(qr).cuda()
    items, scores, counts = base.search_device(qd, k, 0.0, row_to_group=torch.from_numpy(chunk_to_message).cuda())
    torch.cuda.synchronize()
    plain = base.fuzzy_lookup_embeddings(qr, k, 0.0)
    for i in range(b):
        want = fold_chunk_hits_to_messages(plain[i], chunk_to_message)
        c = int(counts[i])
        assert c == len(want)
        assert items[i, :c].tolist() == [m for m, _ in want]
        np.testing.assert_array_equal(scores[i, :c].cpu().numpy(), np.array([s for _, s in want], np.float32))
        assert (items[i, c:] == -1).all()

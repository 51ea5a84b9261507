"""Parity assertions shared by the oracle and CUDA tests.

The bar (BASELINE.json north_star): scores within 1e-4 of the reference and identical
top-k index sets.  Bit-equality of float32 dot products across BLAS kernels / GPU
summation orders is not defined (sgemv order is unspecified), so "identical sets" is
checked up to *exact-arithmetic ties*: any row present on one side only must have a
score within ``tie_tol`` of the rank-k boundary score or of ``min_score`` — i.e. it is a
row whose membership is decided by the last bit of a float32 sum.  Tolerances are
explicit at every call site.
"""

from __future__ import annotations

import numpy as np

SCORE_TOL = 1e-4  # the north_star tolerance on scores
TIE_TOL = 2e-6  # a few float32 ulps at score ~0.5..1.0 (summation-order noise)


def as_pairs(hits):
    """Accept list[ScoredInt|Hit] or dict(items=..., scores=...)."""
    if isinstance(hits, dict):
        return list(hits["items"]), list(hits["scores"])
    return [h.item for h in hits], [h.score for h in hits]


def assert_hits_match(got, want, *, score_tol=SCORE_TOL, tie_tol=TIE_TOL, min_score=None, what=""):
    g_items, g_scores = as_pairs(got)
    w_items, w_scores = as_pairs(want)
    ctx = f"{what}: got {list(zip(g_items, g_scores))[:6]}... want {list(zip(w_items, w_scores))[:6]}..."
    # descending order on both sides
    assert all(a >= b for a, b in zip(g_scores, g_scores[1:])), f"not descending; {ctx}"
    g_map = dict(zip(g_items, g_scores)) if len(set(g_items)) == len(g_items) else None
    if g_map is None:
        # duplicate ordinals (subset lookups may repeat rows): compare as multisets
        assert sorted(g_items) == sorted(w_items), ctx
        np.testing.assert_allclose(sorted(g_scores), sorted(w_scores), atol=score_tol, rtol=0)
        return
    w_map = dict(zip(w_items, w_scores))
    only_g = set(g_map) - set(w_map)
    only_w = set(w_map) - set(g_map)
    if only_g or only_w:
        # boundary = lowest score on either side, or the threshold
        bounds = []
        if g_scores:
            bounds.append(g_scores[-1])
        if w_scores:
            bounds.append(w_scores[-1])
        if min_score is not None:
            bounds.append(float(np.float32(min_score)))
        for r in only_g:
            assert any(abs(g_map[r] - b) <= tie_tol for b in bounds), f"row {r} only in got; {ctx}"
        for r in only_w:
            assert any(abs(w_map[r] - b) <= tie_tol for b in bounds), f"row {r} only in want; {ctx}"
        assert abs(len(g_items) - len(w_items)) <= len(only_g) + len(only_w), ctx
    else:
        assert len(g_items) == len(w_items), ctx
    for r in set(g_map) & set(w_map):
        assert abs(g_map[r] - w_map[r]) <= score_tol, f"row {r}: {g_map[r]} vs {w_map[r]}; {ctx}"
    # order: positions may differ only between near-equal scores
    for pos, (gi, wi) in enumerate(zip(g_items, w_items)):
        if gi != wi:
            assert abs(g_scores[pos] - w_scores[pos]) <= tie_tol, f"order differs at {pos}; {ctx}"


def blocked_oracle_lookup(corpus, queries, k, min_score=0.0, row_offset=0, block_rows=1_000_000):
    """Oracle top-k over a DEVICE-resident corpus too large to score in one piece: the rows come back
    in ``block_rows`` blocks as float32 (storage -> float32 is exact), ``oracle.lookup`` — the
    reference's np.dot / clip / flatnonzero / argpartition — runs per block and the per-block lists
    are merged like shards (exact: top-k of a union of exact per-block top-k lists).  ``corpus`` is a
    torch CUDA tensor [N, D]; ``queries`` float32 numpy [B, D].  Returns one list of Hits per query."""
    import torch

    from oracle import vectorbase_oracle as O

    parts = [[] for _ in range(len(queries))]
    n = corpus.shape[0]
    for lo in range(0, n, block_rows):
        block = corpus[lo:lo + block_rows].to(torch.float32).cpu().numpy()
        for j, q in enumerate(queries):
            hits = O.lookup(block, q, k, min_score)
            parts[j].append([O.Hit(h.item + lo + row_offset, h.score) for h in hits])
        del block
    return [O.merge_shard_hits(p, k) for p in parts]

"""CPU restatement (numpy) of the reference's VectorBase top-k lookup.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): the checker for the CUDA
path and the timed CPU comparator in ``bench.py``; never imported by the product.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks every function
here against (a) outputs of the unmodified reference file generated in the build
container by ``tests/golden/make_golden.py`` (committed under ``tests/golden/``),
(b) the reference's own known-answer tests (tests/test_vectorbase.py:239-252 exact
``[1.0, 0.5, 0.0]``; :209-236 subset cases), and (c) — when /root/reference is
mounted — the live reference on fresh random inputs.

The arithmetic itself lives in a third-party dependency of the reference, numpy
(pyproject.toml:35 ``numpy>=2.2.6``; uv.lock pins 2.4.4; this image has 2.3.x with
OpenBLAS): ``np.dot`` (sgemv, summation order unspecified), ``np.clip``,
``np.flatnonzero``, ``np.argpartition``, ``np.argsort``.  This restatement calls the
same numpy primitives in the same order as the reference so that ties and
selection behave identically; citations are to /root/reference/src/typeagent.

All citations below: aitools/vectorbase.py unless another file is named.
"""

from __future__ import annotations

from collections.abc import Callable, Sequence
from dataclasses import dataclass

import numpy as np

DEFAULT_MIN_SCORE = 0.85  # :16
KNOWN_MODEL_MIN_SCORES = {  # :31-35
    "text-embedding-3-large": 0.74,
    "text-embedding-3-small": 0.73,
    "text-embedding-ada-002": 0.93,
}
DEFAULT_MAX_HITS = 10  # :170-171 (quirk Q1: None means 10, not "unlimited")


@dataclass
class Hit:
    """(row ordinal, score) — the oracle's stand-in for ScoredInt (:50-55)."""

    item: int
    score: float


def score_from_cosine(x: np.ndarray) -> np.ndarray:
    """:44-47 — map a dot product in [-1, 1] onto the public [0, 1] scale, in the
    input's dtype (float32 on the hot path): clip((x + 1) / 2, 0, 1)."""
    return np.clip((x + 1.0) / 2.0, 0.0, 1.0)


def _rank_passing(scores: np.ndarray, k: int, min_score: float):
    """:179-187 / :219-227 — threshold, then top-k.

    Returns (positions, scores) of the selected entries in output order.
    ``scores >= min_score`` compares in float32 (a Python float is a weak scalar
    under NEP 50), which is why the CUDA side takes ``(float)min_score``.
    """
    passing = np.flatnonzero(scores >= min_score)
    if len(passing) == 0:
        return passing, scores[:0]
    kept = scores[passing]
    if len(passing) <= k:
        order = np.argsort(kept)[::-1]
    else:
        part = np.argpartition(kept, -k)[-k:]
        order = part[np.argsort(kept[part])[::-1]]
    return passing[order], kept[order]


def lookup(
    vectors: np.ndarray,
    embedding: np.ndarray,
    max_hits: int | None = None,
    min_score: float | None = None,
    predicate: Callable[[int], bool] | None = None,
) -> list[Hit]:
    """:163-201 ``fuzzy_lookup_embedding``."""
    k = DEFAULT_MAX_HITS if max_hits is None else max_hits
    floor = 0.0 if min_score is None else min_score
    if len(vectors) == 0:  # :174-175
        return []
    scores = score_from_cosine(np.dot(vectors, embedding))  # :176
    if predicate is None:
        rows, vals = _rank_passing(scores, k, floor)
        return [Hit(int(r), float(v)) for r, v in zip(rows, vals)]
    # :191-201 predicate path: threshold -> python filter -> stable sort desc -> [:k]
    hits = [
        Hit(int(r), float(scores[r]))
        for r in np.flatnonzero(scores >= floor)
        if predicate(int(r))
    ]
    hits.sort(key=lambda h: h.score, reverse=True)
    return hits[:k]


def lookup_in_subset(
    vectors: np.ndarray,
    embedding: np.ndarray,
    ordinals_of_subset: Sequence[int],
    max_hits: int | None = None,
    min_score: float | None = None,
) -> list[Hit]:
    """:203-230 ``fuzzy_lookup_embedding_in_subset`` (gather, then the same ranking;
    the returned item is the caller's ordinal, duplicates allowed)."""
    k = DEFAULT_MAX_HITS if max_hits is None else max_hits
    floor = 0.0 if min_score is None else min_score
    if len(ordinals_of_subset) == 0 or len(vectors) == 0:  # :214-215
        return []
    subset = np.asarray(ordinals_of_subset)
    scores = score_from_cosine(np.dot(vectors[subset], embedding))  # :218
    pos, vals = _rank_passing(scores, k, floor)
    return [Hit(int(subset[p]), float(v)) for p, v in zip(pos, vals)]


def lookup_batch(
    vectors: np.ndarray,
    queries: np.ndarray,
    max_hits: int | None = None,
    min_score: float | None = None,
    *,
    one_gemm: bool = False,
) -> list[list[Hit]]:
    """What every reference caller does with several query embeddings
    (storage/memory/reltermsindex.py:320-332): one ``lookup`` per query.

    ``one_gemm=True`` is the "strong CPU baseline" of SURVEY.md §8d: a single
    sgemm ``Q @ V.T`` followed by the per-row ranking (same results up to the
    summation order of the BLAS kernel).
    """
    if not one_gemm:
        return [lookup(vectors, q, max_hits, min_score) for q in queries]
    k = DEFAULT_MAX_HITS if max_hits is None else max_hits
    floor = 0.0 if min_score is None else min_score
    if len(vectors) == 0:
        return [[] for _ in queries]
    all_scores = score_from_cosine(queries @ vectors.T)
    out = []
    for row in all_scores:
        rows, vals = _rank_passing(row, k, floor)
        out.append([Hit(int(r), float(v)) for r, v in zip(rows, vals)])
    return out


def shard_bounds(n_rows: int, n_shards: int) -> list[tuple[int, int]]:
    """Contiguous row blocks, shard g owns [g*ceil(N/G), (g+1)*ceil(N/G))."""
    per = -(-n_rows // n_shards) if n_shards > 0 else 0
    return [(min(g * per, n_rows), min((g + 1) * per, n_rows)) for g in range(n_shards)]


def merge_shard_hits(per_shard: list[list[Hit]], k: int) -> list[Hit]:
    """Top-k of the union of per-shard top-k lists (items already global).  Ties:
    higher score first, then higher row first — the order the CUDA path defines."""
    pool = [h for hits in per_shard for h in hits]
    pool.sort(key=lambda h: (np.float32(h.score), h.item), reverse=True)
    return pool[:k]


def lookup_sharded(
    vectors: np.ndarray,
    embedding: np.ndarray,
    n_shards: int,
    max_hits: int | None = None,
    min_score: float | None = None,
) -> list[Hit]:
    """Row-sharded lookup: the reference lookup on each shard, ordinals shifted to
    global rows, then a k-way merge.  Exact: top-k of a union of exact per-shard
    top-k lists is the global top-k."""
    k = DEFAULT_MAX_HITS if max_hits is None else max_hits
    parts = []
    for lo, hi in shard_bounds(len(vectors), n_shards):
        hits = lookup(vectors[lo:hi], embedding, k, min_score)
        parts.append([Hit(h.item + lo, h.score) for h in hits])
    return merge_shard_hits(parts, k)


# --------------------------------------------------------------------------
# storage-dtype rounding (the bf16 / fp16 configs): the oracle is fed the
# storage-rounded values upcast to float32 ("identical fp32 inputs").
# --------------------------------------------------------------------------


def round_to_bfloat16(x: np.ndarray) -> np.ndarray:
    """float32 -> nearest-even bfloat16 -> float32, bit-exact w.r.t. __float2bfloat16_rn."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    bits = x.view(np.uint32).astype(np.uint64)
    lsb = (bits >> 16) & 1
    rounded = ((bits + 0x7FFF + lsb) >> 16) << 16
    nan = np.isnan(x)
    out = (rounded & 0xFFFFFFFF).astype(np.uint32).view(np.float32).reshape(x.shape)
    if nan.any():
        out = out.copy()
        out[nan] = np.nan
    return out


def round_to_float16(x: np.ndarray) -> np.ndarray:
    return np.asarray(x, dtype=np.float32).astype(np.float16).astype(np.float32)


def round_to_storage(x: np.ndarray, dtype: str) -> np.ndarray:
    if dtype in ("float32", "f32"):
        return np.asarray(x, dtype=np.float32)
    if dtype in ("bfloat16", "bf16"):
        return round_to_bfloat16(x)
    if dtype in ("float16", "f16"):
        return round_to_float16(x)
    raise ValueError(f"unknown storage dtype {dtype!r}")


# --------------------------------------------------------------------------
# synthetic inputs (mirrors tools/benchmark_vectorbase.py:80-94)
# --------------------------------------------------------------------------


def make_corpus(n_rows: int, dim: int, seed: int, n_queries: int = 1):
    """Unit-norm float32 corpus [N, D] and queries [B, D]; queries are drawn from
    the same generator *after* the corpus, as the reference benchmark does."""
    rng = np.random.default_rng(seed)
    vectors = rng.standard_normal((n_rows, dim)).astype(np.float32)
    vectors /= np.linalg.norm(vectors, axis=1, keepdims=True)
    queries = rng.standard_normal((n_queries, dim)).astype(np.float32)
    queries /= np.linalg.norm(queries, axis=1, keepdims=True)
    return vectors, queries


# --------------------------------------------------------------------------
# deterministic fake text embeddings (tests only)
# --------------------------------------------------------------------------


def fake_text_embedding(text: str, dim: int = 3) -> np.ndarray:
    """The reference test model's vector for ``text`` (aitools/model_adapters.py:375-404
    for the hash, :176-184 for the float32 L2 normalisation): component i is the
    31-multiplier polynomial hash (mod 2^32) of ``text`` rotated left by i mod len,
    reduced mod 1961 and divided by 1961."""
    if not text:
        raise ValueError("Empty input text")
    raw = []
    for i in range(dim):
        r = i % len(text)
        acc = 0
        for ch in text[r:] + text[:r]:
            acc = (acc * 31 + ord(ch)) & 0xFFFFFFFF
        raw.append((acc % 1961) / 1961)
    v = np.array([raw], dtype=np.float32)
    norms = np.linalg.norm(v, axis=1, keepdims=True).astype(np.float32)
    norms = np.where(norms > 0, norms, np.float32(1.0))
    return (v / norms).astype(np.float32)[0]


class FakeEmbeddingModel:
    """Caching fake model with the IEmbeddingModel surface (aitools/embeddings.py:39-114)."""

    model_name = "test"

    def __init__(self, dim: int = 3) -> None:
        self._dim = dim
        self._cache: dict[str, np.ndarray] = {}

    def add_embedding(self, key: str, embedding: np.ndarray) -> None:
        self._cache[key] = embedding

    async def get_embedding_nocache(self, input: str) -> np.ndarray:
        return fake_text_embedding(input, self._dim)

    async def get_embeddings_nocache(self, input: list[str]) -> np.ndarray:
        if not input:
            raise ValueError("Cannot embed an empty list")
        return np.stack([fake_text_embedding(t, self._dim) for t in input])

    async def get_embedding(self, key: str) -> np.ndarray:
        hit = self._cache.get(key)
        if hit is None:
            hit = self._cache[key] = fake_text_embedding(key, self._dim)
        return hit

    async def get_embeddings(self, keys: list[str]) -> np.ndarray:
        if not keys:
            raise ValueError("Cannot embed an empty list")
        return np.array([await self.get_embedding(k) for k in keys], dtype=np.float32)


# --------------------------------------------------------------------------
# the whole class, restated (for API-parity tests parametrised over backends)
# --------------------------------------------------------------------------


class OracleVectorBase:
    """numpy restatement of ``VectorBase`` (:82-287) with the same method names,
    argument meaning and error behaviour; ``settings`` needs ``embedding_model``,
    ``min_score`` and ``max_matches`` attributes."""

    def __init__(self, settings) -> None:
        self.settings = settings
        self._model = settings.embedding_model
        self._embedding_size = 0
        self.clear()

    def __len__(self) -> int:
        return len(self._vectors)

    def __bool__(self) -> bool:  # :112-113
        return True

    async def get_embedding(self, key, cache=True):
        m = self._model
        return await (m.get_embedding(key) if cache else m.get_embedding_nocache(key))

    async def get_embeddings(self, keys, cache=True):
        m = self._model
        return await (m.get_embeddings(keys) if cache else m.get_embeddings_nocache(keys))

    def _adopt_width(self, width: int) -> None:
        if self._embedding_size == 0:
            assert width > 0
            self._embedding_size = width
            self._vectors.shape = (0, width)

    def _check_width(self, width: int) -> None:
        if width != self._embedding_size:
            raise ValueError(
                f"Embedding size mismatch: expected {self._embedding_size}, got {width}"
            )

    def add_embedding(self, key, embedding) -> None:  # :115-130
        row = np.asarray(embedding, dtype=np.float32)
        self._adopt_width(len(row))
        self._check_width(len(row))
        self._vectors = np.append(self._vectors, row.reshape(1, -1), axis=0)
        if key is not None:
            self._model.add_embedding(key, row)

    def add_embeddings(self, keys, embeddings) -> None:  # :132-148
        if embeddings.ndim != 2:
            raise ValueError(f"Expected 2D embeddings array, got {embeddings.ndim}D")
        self._adopt_width(embeddings.shape[1])
        self._check_width(embeddings.shape[1])
        self._vectors = np.concatenate((self._vectors, embeddings), axis=0)
        if keys is not None:
            for key, row in zip(keys, embeddings):
                self._model.add_embedding(key, row)

    async def add_key(self, key, cache=True) -> None:  # :150-152
        self.add_embedding(key if cache else None, await self.get_embedding(key, cache=cache))

    async def add_keys(self, keys, cache=True):  # :154-161
        if not keys:
            return None
        rows = await self.get_embeddings(keys, cache=cache)
        self.add_embeddings(keys if cache else None, rows)
        return rows

    def fuzzy_lookup_embedding(self, embedding, max_hits=None, min_score=None, predicate=None):
        return lookup(self._vectors, embedding, max_hits, min_score, predicate)

    def fuzzy_lookup_embedding_in_subset(
        self, embedding, ordinals_of_subset, max_hits=None, min_score=None
    ):
        return lookup_in_subset(self._vectors, embedding, ordinals_of_subset, max_hits, min_score)

    async def fuzzy_lookup(self, key, max_hits=None, min_score=None, predicate=None):  # :232-246
        if max_hits is None:
            max_hits = self.settings.max_matches
        if min_score is None:
            min_score = self.settings.min_score
        return self.fuzzy_lookup_embedding(
            await self.get_embedding(key), max_hits, min_score, predicate
        )

    def clear(self) -> None:  # :253-256
        self._vectors = np.array([], dtype=np.float32)
        if self._embedding_size > 0:
            self._vectors.shape = (0, self._embedding_size)

    def get_embedding_at(self, pos: int):  # :258-263
        if 0 <= pos < len(self._vectors):
            return self._vectors[pos]
        raise IndexError(f"Index {pos} out of bounds for embedding index of size {len(self)}")

    def serialize_embedding_at(self, pos: int):  # :265-266
        return self._vectors[pos] if 0 <= pos < len(self._vectors) else None

    def serialize(self):  # :268-271
        return self._vectors

    def deserialize(self, data) -> None:  # :273-287
        if data is None:
            self.clear()
            return
        if self._embedding_size == 0:
            if data.ndim < 2 or data.shape[0] == 0:
                self.clear()
                return
            self._embedding_size = data.shape[1]
        assert data.shape == (len(data), self._embedding_size), (data.shape, self._embedding_size)
        self._vectors = data

"""``EmbeddingIndex`` over the GPU ``VectorBase`` — the surface of
``typeagent.knowpro.fuzzyindex.EmbeddingIndex`` (/root/reference/src/typeagent/knowpro/
fuzzyindex.py:12-143): push / get / nearest / nearest-in-subset / (de)serialize, plus the
batched ``get_indexes_of_nearest_batch`` that the related-terms expansion can use instead of
its per-term loop (storage/memory/reltermsindex.py:320-332)."""

from __future__ import annotations

from collections.abc import Callable

import numpy as np

from .vectorbase import ScoredInt, TextEmbeddingIndexSettings, VectorBase


class EmbeddingIndex:
    def __init__(
        self,
        settings: TextEmbeddingIndexSettings,
        embeddings: np.ndarray | None = None,
        **vectorbase_options,
    ):
        self._vector_base = VectorBase(settings, **vectorbase_options)
        if embeddings is not None:
            self._vector_base.add_embeddings(None, embeddings)

    def __len__(self) -> int:
        return len(self._vector_base)

    async def size(self) -> int:
        return len(self._vector_base)

    async def is_empty(self) -> bool:
        return len(self._vector_base) == 0

    async def get_embedding(self, key: str, cache: bool = True):
        return await self._vector_base.get_embedding(key, cache)

    def get(self, pos: int):
        return self._vector_base.get_embedding_at(pos)

    def push(self, embeddings: np.ndarray) -> None:
        self._vector_base.add_embeddings(None, embeddings)

    async def add_texts(self, texts: list[str]) -> None:
        await self._vector_base.add_keys(texts)

    def get_indexes_of_nearest(
        self,
        embedding,
        max_matches: int | None = None,
        min_score: float | None = None,
        predicate: Callable[[int], bool] | None = None,
    ) -> list[ScoredInt]:
        return self._vector_base.fuzzy_lookup_embedding(
            embedding, max_hits=max_matches, min_score=min_score, predicate=predicate
        )

    def get_indexes_of_nearest_in_subset(
        self,
        embedding,
        ordinals_of_subset: list[int],
        max_matches: int | None = None,
        min_score: float | None = None,
    ) -> list[ScoredInt]:
        return self._vector_base.fuzzy_lookup_embedding_in_subset(
            embedding, ordinals_of_subset, max_matches, min_score
        )

    def get_indexes_of_nearest_batch(
        self,
        embeddings: np.ndarray,
        max_matches: int | None = None,
        min_score: float | None = None,
    ) -> list[list[ScoredInt]]:
        return self._vector_base.fuzzy_lookup_embeddings(embeddings, max_matches, min_score)

    def clear(self) -> None:
        self._vector_base.clear()

    def serialize(self) -> np.ndarray:
        return self._vector_base.serialize()

    def deserialize(self, embeddings: np.ndarray) -> None:
        # same input contract as the reference (fuzzyindex.py:135-143)
        assert isinstance(embeddings, np.ndarray), type(embeddings)
        assert embeddings.dtype == np.float32, embeddings.dtype
        assert embeddings.ndim == 2, embeddings.shape
        assert (
            self._vector_base._embedding_size == 0
            or embeddings.shape[1] == self._vector_base._embedding_size
        ), embeddings.shape
        self._vector_base.deserialize(embeddings)

"""Embedding data formats either side of the lookup path: how the reference
persists the rows a ``VectorBase`` searches, read straight into the GPU class.

* ``<name>_embeddings.bin`` — raw little-endian float32 ``[relatedCount + messageCount, D]``;
  the split is in ``<name>_data.json`` under ``embeddingFileHeader`` (reference:
  knowpro/serialization.py:35-36 suffixes, :83-98 writer, :183-249 reader; podcasts/podcast.py:
  147-168 ``np.fromfile(...).reshape(-1, embeddingSize)``).  Rows ``[0, relatedCount)`` are the
  related-term vocabulary, the next ``messageCount`` rows the message chunks.
* SQLite BLOBs — one ``ndarray.tobytes()`` float32 vector per row (storage/sqlite/schema.py:
  193-212), loaded with ``np.frombuffer`` and stacked at open (storage/sqlite/messageindex.py:
  33-45, reltermsindex.py:144-156).

Everything here is host-side layout work; the single H2D upload happens at the first lookup.
"""

from __future__ import annotations

import json
import os
from collections.abc import Iterable

import numpy as np

from .vectorbase import TextEmbeddingIndexSettings, VectorBase

DATA_FILE_SUFFIX = "_data.json"
EMBEDDING_FILE_SUFFIX = "_embeddings.bin"


class EmbeddingFormatError(ValueError):
    pass


def read_embedding_file_header(filename_prefix: str) -> dict:
    """``embeddingFileHeader`` of ``<prefix>_data.json``: relatedCount, messageCount, embedding size."""
    with open(filename_prefix + DATA_FILE_SUFFIX, encoding="utf-8") as f:
        data = json.load(f)
    header = data.get("embeddingFileHeader")
    if header is None:
        raise EmbeddingFormatError("Missing embedding file header")
    meta = header.get("modelMetadata") or {}
    return {
        "relatedCount": int(header.get("relatedCount") or 0),
        "messageCount": int(header.get("messageCount") or 0),
        "embeddingSize": int(meta.get("embeddingSize") or 0),
    }


def map_embedding_file(filename_prefix: str, embedding_size: int | None = None, mmap: bool = True):
    """The two row blocks of ``<prefix>_embeddings.bin`` as float32 ``[n, D]`` arrays
    ``(related_terms, messages)`` — memory-mapped views by default (zero host copies; the rows
    are read once, by the H2D upload)."""
    header = read_embedding_file_header(filename_prefix)
    dim = embedding_size or header["embeddingSize"]
    if dim <= 0:
        raise EmbeddingFormatError("embedding size unknown: not in the header and not given")
    path = filename_prefix + EMBEDDING_FILE_SUFFIX
    n_bytes = os.path.getsize(path)
    if n_bytes % (4 * dim) != 0:
        raise EmbeddingFormatError(f"{path}: {n_bytes} bytes is not a whole number of {dim}-float rows")
    rows = n_bytes // (4 * dim)
    if mmap and rows:
        flat = np.memmap(path, dtype="<f4", mode="r", shape=(rows, dim))
    else:
        flat = np.fromfile(path, dtype="<f4").reshape(rows, dim)
    n_rel, n_msg = header["relatedCount"], header["messageCount"]
    if n_rel + n_msg > rows:
        raise EmbeddingFormatError(f"Expected {n_rel + n_msg} embeddings, got {rows}")
    return flat[:n_rel], flat[n_rel : n_rel + n_msg]


def load_embedding_file(filename_prefix: str, settings: TextEmbeddingIndexSettings, **vectorbase_options):
    """``(related_terms_base, message_base)``: two GPU ``VectorBase`` objects over the file's rows."""
    related, messages = map_embedding_file(filename_prefix)
    bases = []
    for block in (related, messages):
        base = VectorBase(settings, **vectorbase_options)
        if len(block):
            base.deserialize(np.ascontiguousarray(block, dtype=np.float32))
        bases.append(base)
    return tuple(bases)


def write_embedding_file(filename_prefix: str, related: np.ndarray | None, messages: np.ndarray | None,
                         extra_json: dict | None = None) -> None:
    """Writer counterpart (the reference's layout, serialization.py:83-98): both blocks appended
    to one ``.bin``, counts and embedding size in the JSON header."""
    blocks = [np.ascontiguousarray(b, dtype="<f4") for b in (related, messages) if b is not None and len(b)]
    dim = blocks[0].shape[1] if blocks else 0
    with open(filename_prefix + EMBEDDING_FILE_SUFFIX, "wb") as f:
        for b in blocks:
            if b.shape[1] != dim:
                raise EmbeddingFormatError("related and message embeddings differ in size")
            b.tofile(f)
    data = dict(extra_json or {})
    data.setdefault("fileHeader", {"version": "0.1"})
    data["embeddingFileHeader"] = {
        "relatedCount": 0 if related is None else len(related),
        "messageCount": 0 if messages is None else len(messages),
        "modelMetadata": {"embeddingSize": dim},
    }
    with open(filename_prefix + DATA_FILE_SUFFIX, "w", encoding="utf-8") as f:
        json.dump(data, f)


def embeddings_from_blobs(blobs: Iterable[bytes | None], embedding_size: int | None = None) -> np.ndarray:
    """SQLite embedding BLOBs -> one float32 ``[n, D]`` array (None blobs skipped, as the reference's
    loaders do); raises when a blob's length disagrees with the others (the reference's embedding-
    size consistency check, storage/sqlite/provider.py:185-226)."""
    raw = [b for b in blobs if b is not None]
    if not raw:
        return np.zeros((0, embedding_size or 0), dtype=np.float32)
    dim = embedding_size or len(raw[0]) // 4
    out = np.empty((len(raw), dim), dtype=np.float32)
    for i, blob in enumerate(raw):
        if len(blob) != 4 * dim:
            raise EmbeddingFormatError(f"Embedding size mismatch: expected {dim}, got {len(blob) // 4}")
        out[i] = np.frombuffer(blob, dtype="<f4")
    return out


def fold_chunk_hits_to_messages(hits, chunk_to_message: np.ndarray | list[int], max_matches: int | None = None):
    """Chunk-level hits -> best score per message, descending (the fold
    storage/memory/messageindex.py:185-207 applies above the lookup): ``hits`` are ScoredInt-like
    (``.item`` = chunk row, ``.score``); returns ``[(message_ordinal, score)]``."""
    best: dict[int, float] = {}
    for h in hits:
        m = int(chunk_to_message[h.item])
        if m not in best or h.score > best[m]:
            best[m] = h.score
    ranked = sorted(best.items(), key=lambda kv: kv[1], reverse=True)
    return ranked if max_matches is None else ranked[:max_matches]

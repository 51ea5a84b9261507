"""On-disk / wire formats either side of the lookup: the reference's
``_embeddings.bin`` + ``_data.json`` pair and SQLite float32 BLOBs (host-side; no GPU needed)."""

from __future__ import annotations

import json
from types import SimpleNamespace

import numpy as np
import pytest

import typeagent_py_b200 as tab
from oracle import vectorbase_oracle as O
from typeagent_py_b200 import formats as F


def test_embedding_file_roundtrip_and_split(tmp_path):
    rng = np.random.default_rng(0)
    related = rng.standard_normal((37, 12)).astype(np.float32)
    messages = rng.standard_normal((5, 12)).astype(np.float32)
    prefix = str(tmp_path / "conv")
    F.write_embedding_file(prefix, related, messages, extra_json={"nameTag": "t"})
    # the layout the reference reads: one raw little-endian float32 blob, header in the JSON
    raw = np.fromfile(prefix + "_embeddings.bin", dtype=np.float32).reshape(-1, 12)
    np.testing.assert_array_equal(raw, np.concatenate([related, messages]))
    hdr = json.load(open(prefix + "_data.json"))["embeddingFileHeader"]
    assert hdr == {"relatedCount": 37, "messageCount": 5, "modelMetadata": {"embeddingSize": 12}}
    got_r, got_m = F.map_embedding_file(prefix)
    np.testing.assert_array_equal(got_r, related)
    np.testing.assert_array_equal(got_m, messages)
    assert isinstance(got_r, np.memmap) or isinstance(got_r.base, np.memmap)  # no host copy
    settings = tab.TextEmbeddingIndexSettings(O.FakeEmbeddingModel())
    rel_base, msg_base = F.load_embedding_file(prefix, settings)
    assert len(rel_base) == 37 and len(msg_base) == 5
    np.testing.assert_array_equal(rel_base.serialize(), related)
    np.testing.assert_array_equal(msg_base.get_embedding_at(4), messages[4])


def test_embedding_file_errors(tmp_path):
    prefix = str(tmp_path / "bad")
    F.write_embedding_file(prefix, np.zeros((3, 4), np.float32), None)
    with open(prefix + "_embeddings.bin", "ab") as f:
        f.write(b"\x00\x00")  # not a whole number of rows
    with pytest.raises(F.EmbeddingFormatError):
        F.map_embedding_file(prefix)
    F.write_embedding_file(prefix, np.zeros((3, 4), np.float32), None)
    data = json.load(open(prefix + "_data.json"))
    data["embeddingFileHeader"]["relatedCount"] = 9
    json.dump(data, open(prefix + "_data.json", "w"))
    with pytest.raises(F.EmbeddingFormatError, match="Expected 9 embeddings"):
        F.map_embedding_file(prefix)
    del data["embeddingFileHeader"]
    json.dump(data, open(prefix + "_data.json", "w"))
    with pytest.raises(F.EmbeddingFormatError, match="Missing embedding file header"):
        F.read_embedding_file_header(prefix)
    with pytest.raises(F.EmbeddingFormatError, match="differ in size"):
        F.write_embedding_file(prefix, np.zeros((1, 4), np.float32), np.zeros((1, 5), np.float32))


def test_sqlite_blobs():
    rows = np.random.default_rng(1).standard_normal((6, 9)).astype(np.float32)
    blobs = [r.tobytes() for r in rows]          # schema.py:198 serialize_embedding
    blobs.insert(2, None)
    np.testing.assert_array_equal(F.embeddings_from_blobs(blobs), rows)
    assert F.embeddings_from_blobs([], 9).shape == (0, 9)
    with pytest.raises(F.EmbeddingFormatError, match="Embedding size mismatch"):
        F.embeddings_from_blobs(blobs + [rows[0, :5].tobytes()])


def test_fold_chunk_hits_to_messages():
    hits = [SimpleNamespace(item=i, score=s) for i, s in [(0, 0.9), (1, 0.95), (2, 0.7), (3, 0.8), (4, 0.1)]]
    chunk_to_msg = [10, 10, 11, 12, 11]
    assert F.fold_chunk_hits_to_messages(hits, chunk_to_msg) == [(10, 0.95), (12, 0.8), (11, 0.7)]
    assert F.fold_chunk_hits_to_messages(hits, chunk_to_msg, 2) == [(10, 0.95), (12, 0.8)]

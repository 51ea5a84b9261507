"""Host logic of the row-sharded lookup on CPU: world_size 2 and 3 over the ``gloo`` backend.

The engine is injected: per-rank search = the CPU oracle on the rank's rows, merge = the
oracle's k-way merge, both speaking the product's packed all-gather layout.  What is under
test is the product code around them (typeagent-py_b200/sharded.py): partitioning, global
ordinals, the single packed all_gather_into_tensor, SPMD append and k clamping.  The CUDA
engine itself (search + tav_merge_topk) is covered by tests/test_gpu_parity.py.
"""

from __future__ import annotations

import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class OracleShardEngine:
    """CPU stand-in with CudaShardEngine's interface (test infrastructure)."""

    def __init__(self):
        self.rows = np.zeros((0, 0), np.float32)

    def n_local(self):
        return len(self.rows)

    def load_rows(self, rows):
        self.rows = np.zeros((0, 0), np.float32) if rows is None else np.array(rows, np.float32)

    def append_rows(self, rows):
        self.rows = rows.copy() if len(self.rows) == 0 else np.concatenate([self.rows, rows])

    def search_packed(self, queries, k, min_score, item_offset):
        from oracle import vectorbase_oracle as O
        from typeagent_py_b200.sharded import packed_layout

        b = len(queries)
        off_s, off_c, total = packed_layout(b, k)
        buf = np.zeros(total, np.uint8)
        items = buf[: b * k * 8].view(np.int64).reshape(b, k)
        scores = buf[off_s: off_s + b * k * 4].view(np.float32).reshape(b, k)
        counts = buf[off_c: off_c + b * 4].view(np.int32)
        items[:] = -1
        for i, q in enumerate(queries):
            hits = O.lookup(self.rows, q, k, min_score) if len(self.rows) else []
            hits.sort(key=lambda h: (np.float32(h.score), h.item), reverse=True)
            counts[i] = len(hits)
            for j, h in enumerate(hits):
                items[i, j], scores[i, j] = h.item + item_offset, h.score
        return torch.from_numpy(buf)

    def merge(self, gathered, world, n_queries, k):
        from oracle import vectorbase_oracle as O
        from typeagent_py_b200.sharded import packed_layout

        off_s, off_c, _ = packed_layout(n_queries, k)
        g = gathered.numpy()
        out_i = np.full((n_queries, k), -1, np.int64)
        out_s = np.zeros((n_queries, k), np.float32)
        out_c = np.zeros(n_queries, np.int32)
        for q in range(n_queries):
            parts = []
            for r in range(world):
                items = g[r, : n_queries * k * 8].view(np.int64).reshape(n_queries, k)
                scores = g[r, off_s: off_s + n_queries * k * 4].view(np.float32).reshape(n_queries, k)
                counts = g[r, off_c: off_c + n_queries * 4].view(np.int32)
                parts.append([O.Hit(int(items[q, j]), float(scores[q, j])) for j in range(counts[q])])
            merged = O.merge_shard_hits(parts, k)
            out_c[q] = len(merged)
            for j, h in enumerate(merged):
                out_i[q, j], out_s[q, j] = h.item, h.score
        return torch.from_numpy(out_i), torch.from_numpy(out_s), torch.from_numpy(out_c)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, n_rows: int):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from types import SimpleNamespace

        from oracle import vectorbase_oracle as O
        from typeagent_py_b200.sharded import ShardedVectorBase, shard_bounds

        v, q = O.make_corpus(n_rows, 24, seed=7, n_queries=5)
        settings = SimpleNamespace(embedding_model=O.FakeEmbeddingModel(), min_score=0.85, max_matches=None)
        sh = ShardedVectorBase(settings, engine=OracleShardEngine())
        assert len(sh) == 0 and sh.fuzzy_lookup_embedding(q[0]) == []
        sh.deserialize(v)
        assert len(sh) == n_rows and sh.local_range == shard_bounds(n_rows, world)[rank]
        for k, ms in ((10, 0.0), (3, 0.5), (n_rows + 9, 0.4)):
            got = sh.fuzzy_lookup_embeddings(q, k, ms)
            for qq, hits in zip(q, got):
                want = O.lookup(v, qq, k, ms)
                assert [h.item for h in hits] == [h.item for h in want], (rank, k, ms)
                np.testing.assert_allclose([h.score for h in hits], [h.score for h in want], atol=2e-6)
        # SPMD append: new rows get the next global ordinals and live on the last rank
        extra = O.make_corpus(17, 24, seed=8)[0]
        sh.add_embeddings(None, extra)
        assert len(sh) == n_rows + 17
        both = np.concatenate([v, extra])
        hit = sh.fuzzy_lookup_embedding(extra[3], 1, 0.0)[0]
        assert hit.item == n_rows + 3 and abs(hit.score - 1.0) < 1e-6
        want = O.lookup(both, q[1], 12, 0.3)
        assert [h.item for h in sh.fuzzy_lookup_embedding(q[1], 12, 0.3)] == [h.item for h in want]
        # load_local_shard: each rank brings only its block
        sh2 = ShardedVectorBase(settings, engine=OracleShardEngine())
        lo, hi = shard_bounds(n_rows, world)[rank]
        sh2.load_local_shard(v[lo:hi], n_rows)
        assert [h.item for h in sh2.fuzzy_lookup_embedding(q[2], 7, 0.0)] == [
            h.item for h in O.lookup(v, q[2], 7, 0.0)]
        with pytest.raises(ValueError):
            sh2.add_embeddings(None, np.zeros((2, 5), np.float32))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_rows", [(2, 1001), (3, 100), (2, 1)])
def test_sharded_lookup_over_gloo(world, n_rows):
    mp.spawn(_worker, args=(world, _free_port(), n_rows), nprocs=world, join=True)


class DeferringOracleEngine(OracleShardEngine):
    """Stand-in for the tensor-core path's deferred exact fallback: a rank in `spoil_ranks` first hands out an
    EMPTY candidate list for query 0 and puts the real one in place at finish() (as tav_finish_search does)."""

    def __init__(self, rank, spoil_ranks, fail_rank=None):
        super().__init__()
        self.rank, self.spoil_ranks, self.fail_rank = rank, spoil_ranks, fail_rank
        self.fixups = []

    def search_packed(self, queries, k, min_score, item_offset, defer_check=False):
        from typeagent_py_b200.sharded import packed_layout

        buf = super().search_packed(queries, k, min_score, item_offset)
        if defer_check and self.rank in self.spoil_ranks:
            _, off_c, _ = packed_layout(len(queries), k)
            counts = buf.numpy()[off_c: off_c + 4 * len(queries)].view(np.int32)
            self.fixups.append((counts, int(counts[0])))
            counts[0] = 0
        return buf

    def finish(self):
        if self.fail_rank == self.rank:
            self.fixups.clear()
            raise RuntimeError("exact fallback failed on this rank")
        n = len(self.fixups)
        for counts, real in self.fixups:
            counts[0] = real
        self.fixups.clear()
        return n


def _deferred_worker(rank: int, world: int, port: int):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from types import SimpleNamespace

        from oracle import vectorbase_oracle as O
        from typeagent_py_b200.sharded import ShardedVectorBase

        v, q = O.make_corpus(301, 16, seed=3, n_queries=4)
        settings = SimpleNamespace(embedding_model=O.FakeEmbeddingModel(), min_score=0.85, max_matches=None)
        sh = ShardedVectorBase(settings, engine=DeferringOracleEngine(rank, spoil_ranks={0}))
        sh.deserialize(v)
        # three deferred searches, ONE finish: every one of them is merged again after the correction
        ks = (5, 9, 3)
        outs = [sh.search_tensors(q, k, 0.0, defer_check=True) for k in ks]
        stale = outs[0][0].numpy().copy()
        assert sh.finish() == 3 and sh.finish() == 0
        for k, (items, scores, counts) in zip(ks, outs):
            for i in range(len(q)):
                want = O.lookup(v, q[i], k, 0.0)
                assert items[i, : counts[i]].tolist() == [h.item for h in want], (rank, k, i)
        lo, hi = sh.local_range
        best0 = O.lookup(v, q[0], 1, 0.0)[0].item
        if 0 <= best0 < shard_hi(301, world, 0):   # the spoiled rank owned query 0's best row: the first merge missed it
            assert stale[0, 0] != best0
        # not deferred: resolved before returning
        items, _, counts = sh.search_tensors(q, 4, 0.0)
        assert items[0, : counts[0]].tolist() == [h.item for h in O.lookup(v, q[0], 4, 0.0)]
        # a rank whose fallback fails must not leave the others in the collective: every rank raises
        bad = ShardedVectorBase(settings, engine=DeferringOracleEngine(rank, spoil_ranks={1}, fail_rank=1))
        bad.deserialize(v)
        bad.search_tensors(q, 5, 0.0, defer_check=True)
        with pytest.raises(RuntimeError):
            bad.finish()
        assert bad.finish() == 0
    finally:
        dist.destroy_process_group()


def shard_hi(n_rows, world, rank):
    from typeagent_py_b200.sharded import shard_bounds

    return shard_bounds(n_rows, world)[rank][1]


def test_deferred_searches_are_all_repaired_at_finish_over_gloo():
    mp.spawn(_deferred_worker, args=(2, _free_port()), nprocs=2, join=True)


def test_shard_bounds_and_packing():
    from typeagent_py_b200.sharded import packed_layout, shard_bounds

    assert shard_bounds(10, 4) == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert shard_bounds(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert shard_bounds(0, 2) == [(0, 0), (0, 0)]
    off_s, off_c, total = packed_layout(3, 5)
    assert off_s == 120 and off_c == 120 + 64 and total == off_c + 16 and total % 8 == 0

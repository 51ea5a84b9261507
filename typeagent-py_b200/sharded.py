"""Row-sharded ``VectorBase`` across the GPUs of one box: one process per GPU
(``torch.distributed``), contiguous row blocks, one candidate all-gather per lookup.

Reference behaviour being scaled out: ``VectorBase.fuzzy_lookup_embedding`` (/root/reference/
src/typeagent/aitools/vectorbase.py:163-201) over the *whole* corpus.  Top-k is a decomposable
reduction, so every rank runs the single-GPU search on its rows (ordinals shifted to global
rows by ``item_offset``) and the per-rank ``[B, k]`` candidate lists — packed into one buffer —
are exchanged and merged on every rank.  Result: identical to the unsharded search, including
tie order.

Two exchanges:
  * ``exchange="peer"`` (default on CUDA): libtavec's own (``tav_sharded_search``, csrc/tav_group.cu) —
    every rank's publish kernel stores its list straight into every peer's HBM over NVLink (CUDA IPC
    mapped exchange regions, system-scope release flags), the merge waits on the flags; no NCCL call,
    no host synchronisation, two tiny launches after the local search.  ``torch.distributed`` only
    carries the 64-byte IPC handles once;
  * ``exchange="nccl"``: ONE ``all_gather_into_tensor`` + ``tav_merge_topk`` (also what the CPU tests
    drive over ``gloo`` with an injected engine).

``torch`` is plumbing here (process group, device buffers); the search, exchange and merge are
libtavec kernels.  The engine is injectable so that the host logic (partitioning, packing, gather,
offsets) is testable on CPU with the ``gloo`` backend.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from .vectorbase import ScoredInt, TextEmbeddingIndexSettings, VectorBase


def shard_bounds(n_rows: int, world: int) -> list[tuple[int, int]]:
    """Rank g owns rows [g*ceil(N/G), (g+1)*ceil(N/G)) clipped to N."""
    per = -(-n_rows // world) if world > 0 else 0
    return [(min(g * per, n_rows), min((g + 1) * per, n_rows)) for g in range(world)]


def packed_layout(n_queries: int, k: int) -> tuple[int, int, int]:
    """Byte offsets (scores, counts) and total size of one rank's packed candidate buffer:
    [items int64 B*k | scores float32 B*k | counts int32 B], each section 8-byte aligned."""
    a8 = lambda v: (v + 7) & ~7  # noqa: E731
    off_scores = a8(n_queries * k * 8)
    off_counts = off_scores + a8(n_queries * k * 4)
    total = off_counts + a8(n_queries * 4)
    return off_scores, off_counts, total


class CudaShardEngine:
    """The product engine: a GPU VectorBase for the local rows + libtavec's merge kernel."""

    def __init__(self, settings, device: int, storage_dtype: str = "float32"):
        import torch

        self.torch = torch
        self.device = torch.device("cuda", device)
        self.base = VectorBase(settings, device=device, storage_dtype=storage_dtype)
        self._group = None        # tav_group handle (peer exchange)
        self._group_keep = []     # outputs of deferred group searches, alive until finish

    # ---- peer exchange (tav_group) --------------------------------------------------------
    GROUP_DEPTH = 8

    def _ensure_group(self, dist, process_group, rank: int, world: int, n_queries: int, k: int):
        """(Re)create this rank's exchange region when the batch shape outgrows it — collectively:
        every rank calls with the same shape, the IPC handles travel through all_gather_object."""
        lib = _capi.load()
        if self._group is not None:
            mq, mk = C.c_int(0), C.c_int(0)
            _capi.check(lib.tav_group_capacity(self._group, C.byref(mq), C.byref(mk), None))
            if n_queries <= mq.value and k <= mk.value:
                return self._group
            self.group_finish()
            self.torch.cuda.synchronize(self.device)
            dist.barrier(group=process_group)       # nobody still publishes into a region about to die
            _capi.check(lib.tav_group_destroy(self._group))
            self._group = None
        handle = C.c_void_p()
        cap_q = max(256, 1 << (max(n_queries, 1) - 1).bit_length())
        cap_k = max(16, 1 << (max(k, 1) - 1).bit_length())
        _capi.check(lib.tav_group_create(self.device.index, rank, world, cap_q, cap_k, self.GROUP_DEPTH,
                                         C.byref(handle)))
        nbytes = lib.tav_group_handle_bytes()
        mine = C.create_string_buffer(nbytes)
        _capi.check(lib.tav_group_local_handle(handle, mine))
        gathered = [None] * world
        dist.all_gather_object(gathered, bytes(mine.raw), group=process_group)
        _capi.check(lib.tav_group_connect(handle, b"".join(gathered)))
        dist.barrier(group=process_group)
        self._group = handle
        return handle

    def group_search(self, dist, process_group, rank, world, queries, k, min_score, item_offset, defer_check):
        torch = self.torch
        if isinstance(queries, np.ndarray):
            queries = torch.from_numpy(np.ascontiguousarray(queries, dtype=np.float32)).to(self.device, non_blocking=True)
        b = queries.shape[0]
        group = self._ensure_group(dist, process_group, rank, world, b, k)
        lib, ix = self.base._ensure_device()
        items = torch.empty((b, k), dtype=torch.int64, device=self.device)
        scores = torch.empty((b, k), dtype=torch.float32, device=self.device)
        counts = torch.empty((b,), dtype=torch.int32, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        flags = self.base._flags() | (_capi.TAV_DEFER_RETRY if defer_check else 0)
        _capi.check(lib.tav_sharded_search(ix, group, C.c_void_p(queries.data_ptr()), b, k, C.c_float(min_score), flags,
                                           item_offset, C.c_void_p(items.data_ptr()), C.c_void_p(scores.data_ptr()),
                                           C.c_void_p(counts.data_ptr()), C.c_void_p(stream)))
        if defer_check:
            self._group_keep.append((queries, items, scores, counts, stream))
        return items, scores, counts

    def group_finish(self) -> int:
        if self._group is None or not self._group_keep:
            return 0
        stream = self._group_keep[-1][4]
        redone = C.c_int(0)
        try:
            _capi.check(_capi.load().tav_sharded_finish(self.base._ix, self._group, C.c_void_p(stream), C.byref(redone)))
        finally:
            self._group_keep.clear()
        return redone.value

    def __del__(self):
        try:
            if self._group is not None:
                _capi.load().tav_group_destroy(self._group)
                self._group = None
        except Exception:
            pass

    def comm_device(self):
        return self.device

    def n_local(self) -> int:
        return len(self.base)

    def load_rows(self, rows: np.ndarray | None) -> None:
        self.base.clear()
        if rows is not None and len(rows):
            self.base.add_embeddings(None, np.ascontiguousarray(rows, dtype=np.float32))

    def adopt_tensor(self, tensor) -> None:
        self.base = VectorBase.from_device_tensor(self.base.settings, tensor)

    def append_rows(self, rows: np.ndarray) -> None:
        self.base.add_embeddings(None, rows)

    def finish(self) -> int:
        return self.base.finish_search()

    def search_packed(self, queries, k: int, min_score: float, item_offset: int, defer_check: bool = False):
        torch = self.torch
        if isinstance(queries, np.ndarray):
            queries = torch.from_numpy(np.ascontiguousarray(queries, dtype=np.float32)).to(
                self.device, non_blocking=True)
        b = queries.shape[0]
        off_s, off_c, total = packed_layout(b, k)
        buf = torch.empty(total, dtype=torch.uint8, device=self.device)
        items = buf[: b * k * 8].view(torch.int64).view(b, k)
        scores = buf[off_s : off_s + b * k * 4].view(torch.float32).view(b, k)
        counts = buf[off_c : off_c + b * 4].view(torch.int32)
        if self.n_local() == 0:
            counts.zero_()
        else:
            self.base.search_device(queries, k, min_score, item_offset=item_offset,
                                    out=(items, scores, counts), defer_check=defer_check)
        return buf

    def merge(self, gathered, world: int, n_queries: int, k: int):
        """gathered: uint8 [world, total] on the device -> (items, scores, counts) tensors."""
        torch = self.torch
        off_s, off_c, total = packed_layout(n_queries, k)
        items = torch.empty((n_queries, k), dtype=torch.int64, device=self.device)
        scores = torch.empty((n_queries, k), dtype=torch.float32, device=self.device)
        counts = torch.empty((n_queries,), dtype=torch.int32, device=self.device)
        base = gathered.data_ptr()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        lib = _capi.load()
        _capi.check(
            lib.tav_merge_topk(self.device.index, world, n_queries, k, C.c_void_p(base),
                               C.c_void_p(base + off_s), C.c_void_p(base + off_c),
                               total // 8, total // 4, total // 4,
                               C.c_void_p(items.data_ptr()), C.c_void_p(scores.data_ptr()),
                               C.c_void_p(counts.data_ptr()), C.c_void_p(stream))
        )
        return items, scores, counts


class ShardedVectorBase:
    """Global corpus [N, D] partitioned by contiguous row blocks over the ranks of a
    process group.  Every rank calls every method with the same arguments (SPMD)."""

    def __init__(self, settings: TextEmbeddingIndexSettings, *, process_group=None,
                 device: int | None = None, storage_dtype: str = "float32", engine=None,
                 exchange: str = "peer"):
        import torch.distributed as dist

        if exchange not in ("peer", "nccl"):
            raise ValueError("exchange must be 'peer' or 'nccl'")
        self.exchange = exchange
        self.settings = settings
        self._dist = dist
        self._group = process_group
        self.rank = dist.get_rank(process_group)
        self.world = dist.get_world_size(process_group)
        if engine is None:
            import os

            if device is None:
                device = int(os.environ.get("LOCAL_RANK", self.rank))
            engine = CudaShardEngine(settings, device, storage_dtype)
        self._engine = engine
        self._starts = [0] * (self.world + 1)  # global row where each rank's block starts
        self._embedding_size = 0
        self._pending: list = []  # deferred searches since the last finish(): ["group"] or (local, b, k, out) tuples

    # ---- corpus ------------------------------------------------------------------------
    def __len__(self) -> int:
        return self._starts[-1]

    def __bool__(self) -> bool:
        return True

    @property
    def local_range(self) -> tuple[int, int]:
        return self._starts[self.rank], self._starts[self.rank + 1]

    def _set_bounds(self, bounds) -> None:
        self._starts = [lo for lo, _ in bounds] + [bounds[-1][1]]

    def deserialize(self, data: np.ndarray | None) -> None:
        """Bulk load: every rank passes the same global float32 [N, D] array (or a
        memory-map of it) and keeps only its own block."""
        if data is None or data.ndim < 2 or len(data) == 0:
            self._engine.load_rows(None)
            self._starts = [0] * (self.world + 1)
            return
        self._embedding_size = data.shape[1]
        bounds = shard_bounds(len(data), self.world)
        self._set_bounds(bounds)
        lo, hi = bounds[self.rank]
        self._engine.load_rows(data[lo:hi])

    def load_local_shard(self, rows, global_rows: int) -> None:
        """Each rank supplies only its own block (numpy rows or a CUDA tensor) of a global
        corpus of ``global_rows`` rows partitioned by ``shard_bounds``."""
        bounds = shard_bounds(global_rows, self.world)
        lo, hi = bounds[self.rank]
        if len(rows) != hi - lo:
            raise ValueError(f"rank {self.rank} must hold rows [{lo}, {hi}), got {len(rows)} rows")
        self._set_bounds(bounds)
        self._embedding_size = rows.shape[1]
        if isinstance(rows, np.ndarray):
            self._engine.load_rows(rows)
        else:
            self._engine.adopt_tensor(rows)

    def add_embeddings(self, keys, embeddings: np.ndarray) -> None:
        """Append rows (global ordinals continue at len(self)); they join the LAST rank's
        block so that blocks stay contiguous and ordered.  ``rebalance`` evens blocks out."""
        if embeddings.ndim != 2:
            raise ValueError(f"Expected 2D embeddings array, got {embeddings.ndim}D")
        if self._embedding_size == 0:
            self._embedding_size = embeddings.shape[1]
        if embeddings.shape[1] != self._embedding_size:
            raise ValueError(
                f"Embedding size mismatch: expected {self._embedding_size}, got {embeddings.shape[1]}")
        if self.rank == self.world - 1:
            self._engine.append_rows(np.ascontiguousarray(embeddings, dtype=np.float32))
        self._starts[-1] += len(embeddings)
        if keys is not None:
            for key, row in zip(keys, embeddings):
                self.settings.embedding_model.add_embedding(key, row)

    # ---- lookups -----------------------------------------------------------------------
    def _gather_and_merge(self, local, b: int, k: int):
        import torch

        if self.world == 1:
            gathered = local.view(1, -1)
        else:
            gathered = torch.empty((self.world, local.numel()), dtype=torch.uint8, device=local.device)
            self._dist.all_gather_into_tensor(gathered.view(-1), local, group=self._group)
        return self._engine.merge(gathered, self.world, b, k)

    def search_tensors(self, queries, k: int, min_score: float = 0.0, defer_check: bool = False):
        """SPMD lookup; returns engine tensors (items, scores, counts), replicated on every
        rank.  ``queries``: float32 [B, D] numpy array or engine-device tensor.

        The local search, the candidate all-gather and the merge are enqueued back to back
        without a host synchronisation; the (rare) "redo this query exactly" check runs at the
        end — immediately, or in ``finish()`` when ``defer_check`` is set — and repeats the
        exchange only if some rank actually had to redo a query."""
        n = len(self)
        b = int(queries.shape[0])
        k = max(1, min(int(k), max(n, 1)))
        lo, _ = self.local_range
        if self.exchange == "peer" and self.world > 1 and hasattr(self._engine, "group_search"):
            out = self._engine.group_search(self._dist, self._group, self.rank, self.world, queries, k,
                                            float(np.float32(min_score)), lo, defer_check)
            self._pending = ["group"] if defer_check else []
            return out
        deferrable = hasattr(self._engine, "finish")
        local = (self._engine.search_packed(queries, k, float(np.float32(min_score)), lo, defer_check=True)
                 if deferrable else self._engine.search_packed(queries, k, float(np.float32(min_score)), lo))
        out = self._gather_and_merge(local, b, k)
        pending = getattr(self, "_pending", None) or []
        if pending == ["group"]:
            pending = []
        if deferrable:
            pending.append((local, b, k, out))
        self._pending = pending
        if not defer_check:
            self.finish()
        return out

    def finish(self) -> int:
        """Resolve a deferred lookup on every rank; returns the number of queries (summed over
        ranks) that took the exact fallback.  Collective: every rank must call it."""
        import torch

        pending = getattr(self, "_pending", None)
        if not pending:
            return 0
        self._pending = []
        if pending == ["group"]:       # libtavec agrees across ranks inside tav_sharded_finish
            return self._engine.group_finish()
        # a local failure must not leave the other ranks waiting in the collective: reduce an error
        # flag together with the count and raise on every rank
        error = None
        try:
            redone = self._engine.finish()
        except Exception as e:  # noqa: BLE001
            error, redone = e, 0
        total, failed = redone, int(error is not None)
        if self.world > 1:
            t = torch.tensor([redone, failed], dtype=torch.int32, device=pending[-1][0].device)
            self._dist.all_reduce(t, group=self._group)
            total, failed = int(t[0].item()), int(t[1].item())
        if failed:
            raise error if error is not None else RuntimeError("finish(): another rank failed its exact fallback")
        if total > 0:  # some shard corrected its candidates (in place): exchange and merge every open search again
            for local, b, k, out in pending:
                items, scores, counts = self._gather_and_merge(local, b, k)
                out[0].copy_(items), out[1].copy_(scores), out[2].copy_(counts)
        return total

    def search_arrays(self, queries: np.ndarray, k: int, min_score: float = 0.0):
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q.reshape(1, -1)
        if q.shape[1] != self._embedding_size and len(self):
            raise ValueError("query width does not match the embedding size")
        if len(self) == 0 or np.isnan(np.float32(min_score)):
            return (np.full((len(q), 1), -1, np.int64), np.zeros((len(q), 1), np.float32),
                    np.zeros(len(q), np.int32))
        items, scores, counts = self.search_tensors(q, k, min_score)
        return items.cpu().numpy(), scores.cpu().numpy(), counts.cpu().numpy()

    def fuzzy_lookup_embeddings(self, embeddings, max_hits=None, min_score=None):
        if min_score is None:
            min_score = 0.0
        if len(self) == 0:
            return [[] for _ in range(len(embeddings))]
        k = VectorBase._resolve_k(max_hits, len(self))
        items, scores, counts = self.search_arrays(embeddings, k, min_score)
        il, sl, cl = items.tolist(), scores.tolist(), counts.tolist()
        return [[ScoredInt(i, s) for i, s in zip(il[b][:c], sl[b][:c])] for b, c in enumerate(cl)]

    def fuzzy_lookup_embedding(self, embedding, max_hits=None, min_score=None):
        return self.fuzzy_lookup_embeddings(np.asarray(embedding, np.float32).reshape(1, -1),
                                            max_hits, min_score)[0]

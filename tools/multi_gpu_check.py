#!/usr/bin/env python3
"""Multi-GPU parity check, launched one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29517 tools/multi_gpu_check.py

Every rank builds the same seeded corpus on the host, keeps its own row block on its GPU
(ShardedVectorBase), and checks that the sharded lookup — local search, candidate exchange (libtavec's
peer-memory publish/merge over NVLink, and the NCCL all-gather form), merge kernel — is BIT-IDENTICAL to
the single-GPU lookup over the whole corpus on that rank's GPU, for the float32 row-scan path, the
float32 split form and the bf16 / fp16 tensor-core paths, and agrees with the CPU oracle.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import typeagent_py_b200 as tab  # noqa: E402
from oracle import vectorbase_oracle as O  # noqa: E402
from tests.parity import assert_hits_match  # noqa: E402
from typeagent_py_b200.sharded import ShardedVectorBase  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    settings = tab.TextEmbeddingIndexSettings(O.FakeEmbeddingModel())
    cases = (("float32", 30011, 96, 7, 25, 0.45), ("bfloat16", 200003, 256, 200, 100, 0.0),
             ("float16", 70001, 128, 33, 10, 0.5), ("float32", 90001, 128, 40, 20, 0.0),
             ("bfloat16", 120000, 64, 300, 5, 0.0), ("float16", 150000, 128, 1024 + 5, 100, 0.0))
    for ci, (storage, n, d, b, k, ms) in enumerate(cases):
        exchange = "nccl" if ci == 2 else "peer"
        v, q = O.make_corpus(n, d, seed=n, n_queries=b)
        vr, qr = O.round_to_storage(v, storage), O.round_to_storage(q, storage)
        sh = ShardedVectorBase(settings, device=local, storage_dtype=storage, exchange=exchange)
        sh.deserialize(v)
        whole = tab.VectorBase(settings, device=local, storage_dtype=storage)
        whole.add_embeddings(None, v)
        got = sh.search_arrays(qr, k, ms)
        want = whole.search_arrays(qr, k, ms)
        assert sh._engine.base.last_timing()["path"] == whole.last_timing()["path"]
        np.testing.assert_array_equal(got[2], want[2])
        for i in range(b):
            c = want[2][i]
            np.testing.assert_array_equal(got[0][i, :c], want[0][i, :c])
            np.testing.assert_array_equal(got[1][i, :c], want[1][i, :c])
        for i in range(min(b, 5)):
            assert_hits_match({"items": got[0][i, : got[2][i]].tolist(), "scores": got[1][i, : got[2][i]].tolist()},
                              O.lookup(vr, qr[i], k, ms), min_score=ms)
        # pipelined: several deferred searches (different queries each), one finish
        qd = [torch.from_numpy(np.ascontiguousarray(np.roll(qr, j, axis=0))).cuda() for j in range(3)]
        outs = [sh.search_tensors(x, k, ms, defer_check=True) for x in qd]
        assert sh.finish() == 0
        torch.cuda.synchronize()
        for j, (it, sc, ct) in enumerate(outs):
            np.testing.assert_array_equal(np.roll(ct.cpu().numpy(), -j, axis=0), want[2])
            np.testing.assert_array_equal(np.roll(it.cpu().numpy(), -j, axis=0)[:, :1], want[0][:, :1])
        # append goes to the last rank and is found globally
        sh.add_embeddings(None, qr[:3])
        hit = sh.fuzzy_lookup_embedding(qr[1], 1, 0.0)[0]
        assert hit.item == n + 1, hit
        dist.barrier()
        if rank == 0:
            print(f"multi-gpu ok: world={world} {storage} n={n} d={d} b={b} k={k} path={whole.last_timing()['path']} "
                  f"exchange={exchange}", flush=True)
    # pathological scores (all rows identical): every rank's tensor-core search flags its queries,
    # finish() redoes them exactly and repeats the exchange
    row = O.round_to_bfloat16(O.make_corpus(1, 64, seed=9)[0])
    same = np.repeat(row, 20000 * world, axis=0)
    for exchange in ("peer", "nccl"):
        sh = ShardedVectorBase(settings, device=local, storage_dtype="bfloat16", exchange=exchange)
        sh.deserialize(same)
        nq = 16   # enough queries for the tensor-core path (fewer go to the exact row scan directly)
        items, scores, counts = sh.search_arrays(np.repeat(row, nq, axis=0), 6, 0.0)
        n = len(same)
        assert sh._engine.base.last_timing()["path"] == "mma"
        assert items.tolist() == [list(range(n - 1, n - 7, -1))] * nq, items
        # three deferred searches of the same pathological corpus, ONE finish: every one of them is repaired
        qd = torch.from_numpy(np.repeat(row, nq, axis=0)).cuda()
        ks = (6, 4, 9)
        outs = [sh.search_tensors(qd, kk, 0.0, defer_check=True) for kk in ks]
        assert sh.finish() > 0
        torch.cuda.synchronize()
        for kk, (it, sc, ct) in zip(ks, outs):
            assert ct.cpu().tolist() == [kk] * nq, ct
            assert it.cpu().tolist() == [list(range(n - 1, n - 1 - kk, -1))] * nq, (kk, it)
        if rank == 0:
            print(f"multi-gpu ok: world={world} exact fallback through finish(), one and three outstanding, "
                  f"exchange={exchange}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

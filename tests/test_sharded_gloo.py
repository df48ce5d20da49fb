"""The N>1 host path on CPU: two gloo ranks shard the corpus by rows, exchange their local
top-k with an all-gather and merge on (score desc, chunk id asc). The local top-k lists come
from the oracle (this is a test of the partition + exchange + merge protocol, not of the
kernels): the merged result must equal the oracle's top-k over the whole corpus, bit for bit,
for every world size — the property the multi-GPU path relies on."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank: int, world: int, port: int, ret):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import oracle
    from rag_era_b200.sharded import allgather_bytes, broadcast_unique_id, merge_reference_order, shard_range
    from rag_era_b200.index import VectorIndex

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        total, d, k = 3001, 64, 10
        g = oracle.make_gen(total, n_clusters=8, dup_period=5)       # exact ties across the shard boundary
        base, n = shard_range(total, world, rank)
        X = oracle.gen_rows(g, base, n, d, threads=1)
        q = oracle.gen_queries(g, 0, 4, d)
        loc_ids = np.full((4, k), np.iinfo(np.uint64).max, dtype=np.uint64)
        loc_sc = np.full((4, k), -np.inf)
        for b in range(4):
            i, s = oracle.topk(X, q[b], k, id_base=base, threads=1)
            loc_ids[b, :len(i)], loc_sc[b, :len(s)] = i, s
        ids_t, sc_t = torch.from_numpy(loc_ids.view(np.int64)), torch.from_numpy(loc_sc)
        all_ids = [torch.empty_like(ids_t) for _ in range(world)]
        all_sc = [torch.empty_like(sc_t) for _ in range(world)]
        dist.all_gather(all_ids, ids_t)
        dist.all_gather(all_sc, sc_t)
        # one binary sidecar, every rank reads (and verifies) only its own row range — what open_sharded_cache streams to HBM
        from rag_era_b200 import _native as N
        cache = os.path.join(os.environ["RAGERA_TEST_TMP"], "kb.ragera")
        if rank == 0:
            N.cache_write_host(cache, oracle.gen_rows(g, 0, total, d, threads=1), ids=[f"n{i}" for i in range(total)])
        dist.barrier()
        part = N.cache_read_host(cache, first_row=base, nrows=n)
        shard_ok = np.array_equal(part["rows"], X) and part["info"].rows == total and len(part["ids"]) == total
        # the unique-id broadcast used for the NCCL communicator (id generation stubbed: no GPU here)
        VectorIndex.comm_unique_id = staticmethod(lambda: bytes(range(128)))
        uid = broadcast_unique_id(dist, rank)
        ok = uid == bytes(range(128)) and shard_ok
        # the host channel of the mailbox bootstrap: every rank's 64-byte handle, ordered by rank
        hs = allgather_bytes(dist, bytes([rank]) * 64)
        ok = ok and hs == [bytes([r]) * 64 for r in range(world)]
        if rank == 0:
            Xall = oracle.gen_rows(g, 0, total, d, threads=1)
            for b in range(4):
                mi, ms = merge_reference_order([a[b].numpy().view(np.uint64) for a in all_ids],
                                               [a[b].numpy() for a in all_sc], k)
                ei, es = oracle.topk(Xall, q[b], k, threads=1)
                ok = ok and np.array_equal(mi, ei) and np.array_equal(ms, es)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_two_rank_shard_exchange_merge(world):
    import torch.multiprocessing as mp

    import tempfile

    tmp = tempfile.mkdtemp(prefix="ragera_gloo_")
    os.environ["RAGERA_TEST_TMP"] = tmp          # inherited by the spawned ranks
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    import shutil

    shutil.rmtree(tmp, ignore_errors=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)

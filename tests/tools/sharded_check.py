#!/usr/bin/env python3
"""Multi-GPU parity check (run under torchrun, one rank per GPU):
row-sharded hybrid search (peer-to-peer mailbox exchange, or RAGERA_COMM=nccl) must equal the CPU oracle over the WHOLE corpus —
ids, scores, fused order and source flags bit for bit — on the stream, tensor and exact paths,
including exact ties that straddle the shard boundary and queries that must escalate.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/tools/sharded_check.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import oracle
    import rag_era_b200 as rb
    from rag_era_b200 import _native as N
    from rag_era_b200.sharded import create_sharded_index, leave_exchange, shard_range

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for (total, d, dt, shadow, dup) in [(200_001, 256, N.F32, True, 7), (150_000, 512, N.BF16, False, 0)]:
        go = oracle.make_gen(total, n_clusters=64, dup_period=dup, memory_rows=total // 3)
        gn = N.GenDesc.from_buffer_copy(bytes(go))
        idx = create_sharded_index(dist, total, d, dt, local, bf16_shadow=shadow)
        base, n = shard_range(total, world, rank)
        idx.generate(gn, n)
        B = 40
        Q = idx.generate_queries(gn, 0, B)
        X = oracle.gen_rows(go, 0, total, d, dtype=oracle.F32 if dt == N.F32 else oracle.BF16) if rank == 0 else None
        rng = np.random.default_rng(5)
        kw = [rng.integers(0, total, 8).tolist() for _ in range(B)]
        for path in (N.PATH_STREAM, N.PATH_TENSOR, N.PATH_EXACT):
            res = idx.hybrid(Q, rb.hybrid_opts(10, 8, 0.3, path=path), kw)
            top = idx.query(Q, 23, path=path)
            if rank == 0:
                for b in range(B):
                    e = oracle.hybrid_search(X, Q[b], 10, 0.3, kw[b])
                    g = res.row(b)
                    good = (np.array_equal(g["keys"], e["keys"]) and np.array_equal(g["scores"], e["scores"]) and
                            np.array_equal(g["source"], e["source"]) and np.array_equal(g["vec_ids"], e["vec_ids"]) and
                            np.array_equal(g["vec_scores"], e["vec_scores"]) and g["certified"])
                    ei, es = oracle.topk(X, Q[b], 23)
                    good = good and np.array_equal(top.row(b)[0], ei) and np.array_equal(top.row(b)[1], es)
                    if not good:
                        ok = False
                        print(f"MISMATCH rows={total} path={path} query={b}", flush=True)
        # forced escalation is collective: every rank must take the same decisions
        r = idx.query(Q[:8], 10, path=N.PATH_STREAM, epsilon=10.0)
        if rank == 0:
            for b in range(8):
                ei, es = oracle.topk(X, Q[b], 10)
                if not (np.array_equal(r.row(b)[0], ei) and np.array_equal(r.row(b)[1], es) and r.certified[b]):
                    ok = False
                    print(f"MISMATCH (escalation) rows={total} query={b}", flush=True)
        leave_exchange(dist, idx)
        idx.close()
    # one binary sidecar, every rank streams only its own row range into HBM (open_sharded_cache)
    import tempfile
    from rag_era_b200.sharded import open_sharded_cache

    total, d = 30_011, 128
    go = oracle.make_gen(total, n_clusters=16)
    path = os.path.join(tempfile.gettempdir(), "ragera_sharded_check.ragera")
    Xc = oracle.gen_rows(go, 0, total, d)
    if rank == 0:
        N.cache_write_host(path, Xc, ids=[f"n{i}" for i in range(total)])
    dist.barrier()
    idx, ids = open_sharded_cache(dist, path, local, bf16_shadow=True)
    Qc = oracle.gen_queries(go, 0, 12, d)
    for pth in (N.PATH_STREAM, N.PATH_TENSOR):
        top = idx.query(Qc, 10, path=pth)
        if rank == 0:
            for b in range(12):
                ei, es = oracle.topk(Xc, Qc[b], 10)
                if not (np.array_equal(top.row(b)[0], ei) and np.array_equal(top.row(b)[1], es) and len(ids) == total):
                    ok = False
                    print(f"MISMATCH (sidecar shards) path={pth} query={b}", flush=True)
    leave_exchange(dist, idx)
    idx.close()
    dist.barrier()
    if rank == 0:
        os.remove(path)
    t = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("sharded parity OK" if int(t.item()) == 1 else "sharded parity FAILED", f"(world={world})", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Row-sharded search with SEVERAL RANKS ON ONE GPU (one process per rank, all on device 0): the whole multi-rank
data path — id_base shards, exact local top-k, the peer-to-peer mailbox exchange fused into K5 / the small-batch
K3+K4+K5 kernel, cross-rank merge, collective escalation — checked against the oracle over the WHOLE corpus on a
single-GPU box. The mailbox handles travel through files in a scratch directory (the bootstrap is host-driven: no
NCCL, no torch.distributed). Also: a rank that skips a call makes its peer's call fail with RAG_ERR_TIMEOUT
(no hang, no trap), the communicator refuses further use, and a fresh bootstrap brings it back.

    python tests/tools/ranks_one_gpu.py <rank> <nranks> <scratch dir>      (tests/test_gpu_sharded.py spawns these)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


class Rendezvous:
    """all-gather of small byte strings through files (atomic rename), ordered by rank"""

    def __init__(self, d, rank, world):
        self.d, self.rank, self.world, self.n = d, rank, world, 0

    def allgather(self, payload: bytes, timeout=240.0):
        self.n += 1
        mine = os.path.join(self.d, f"x{self.n}.{self.rank}")
        with open(mine + ".tmp", "wb") as f:
            f.write(payload)
        os.replace(mine + ".tmp", mine)
        out, t0 = [], time.time()
        for r in range(self.world):
            p = os.path.join(self.d, f"x{self.n}.{r}")
            while not os.path.exists(p):
                if time.time() - t0 > timeout:
                    raise TimeoutError(f"rank {r} never reached rendezvous {self.n}")
                time.sleep(0.01)
            out.append(open(p, "rb").read())
        return out

    def barrier(self):
        self.allgather(b"b")


def main():
    rank, world, scratch = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    import oracle
    import rag_era_b200 as rb
    from rag_era_b200 import _native as N
    from rag_era_b200.sharded import shard_range

    rv = Rendezvous(scratch, rank, world)
    ok = True

    def bad(msg):
        nonlocal ok
        ok = False
        print(f"[rank {rank}] MISMATCH {msg}", flush=True)

    def bootstrap(idx, max_batch, max_k, timeout_ms, step0=0):
        os.environ["RAGERA_P2P_TIMEOUT_MS"] = str(timeout_ms)          # read when the communicator is created
        os.environ["RAGERA_P2P_STEP0"] = str(step0)                     # diagnostics: where the exchange counter starts
        idx.comm_p2p_import(rv.allgather(idx.comm_p2p_export(world, rank, max_batch, max_k)))

    def leave(idx):
        idx.comm_detach()                                               # unmap the peers, barrier, then free
        rv.barrier()
        idx.comm_destroy()

    for (total, d, dt, shadow, dup) in [(60_001, 256, N.F32, "f16", 7), (50_000, 512, N.BF16, None, 0)]:
        go = oracle.make_gen(total, n_clusters=32, dup_period=dup, memory_rows=total // 3)   # dup: ties across the shard boundary
        gn = N.GenDesc.from_buffer_copy(bytes(go))
        base, n = shard_range(total, world, rank)
        idx = rb.VectorIndex(d, n, dtype=dt, device=0, shadow=shadow, id_base=base)
        idx.generate(gn, n)
        # ranks time-slice ONE GPU and check on the CPU in between: be patient. The second corpus starts its exchange counter
        # 5 exchanges before the 20-bit wrap: everything below then runs across it (the mailbox halves must keep alternating)
        bootstrap(idx, 128, 32, 120_000, step0=0 if dt == N.F32 else 0xFFFFA)
        B = 72
        Q = idx.generate_queries(gn, 0, B)
        X = oracle.gen_rows(go, 0, total, d, dtype=oracle.F32 if dt == N.F32 else oracle.BF16)
        rng = np.random.default_rng(5)
        kw = [rng.integers(0, total, 8).tolist() for _ in range(B)]

        def check_hybrid(res, qs, tag):
            for j, b in enumerate(qs):
                e = oracle.hybrid_search(X, Q[b], 10, 0.3, kw[b])
                g = res.row(j)
                if not (np.array_equal(g["keys"], e["keys"]) and np.array_equal(g["scores"], e["scores"]) and
                        np.array_equal(g["source"], e["source"]) and np.array_equal(g["vec_ids"], e["vec_ids"]) and
                        np.array_equal(g["vec_scores"], e["vec_scores"]) and g["certified"]):
                    bad(f"{tag} rows={total} query={b}")

        # batch 1 and small batches: K1 / K1m -> K3+K4+exchange+K5 in ONE kernel
        for b in range(3):
            check_hybrid(idx.hybrid(Q[b:b + 1], rb.hybrid_opts(10, 8, 0.3, path=N.PATH_STREAM), [kw[b]]), [b], "batch-1 stream")
        check_hybrid(idx.hybrid(Q[:9], rb.hybrid_opts(10, 8, 0.3, path=N.PATH_STREAM), kw[:9]), range(9), "batch-9 stream")
        if shadow == "f16":
            # one query streamed from the fp16 shadow of every rank's shard (explicitly and as AUTO's choice), same fused tail
            for b in range(3):
                check_hybrid(idx.hybrid(Q[b:b + 1], rb.hybrid_opts(10, 8, 0.3, path=N.PATH_SHADOW_STREAM), [kw[b]]), [b], "batch-1 shadow stream")
                check_hybrid(idx.hybrid(Q[b:b + 1], rb.hybrid_opts(10, 8, 0.3), [kw[b]]), [b], "batch-1 auto")   # small shard: the fp32 stream
            check_hybrid(idx.hybrid(Q[:5], rb.hybrid_opts(10, 8, 0.3, path=N.PATH_SHADOW_STREAM), kw[:5]), range(5), "batch-5 shadow stream")
        check_hybrid(idx.hybrid(Q[:16], rb.hybrid_opts(10, 8, 0.3, path=N.PATH_TENSOR), kw[:16]), range(16), "batch-16 tensor")
        # larger batches: K3, K4 and K5 (with the exchange) as separate launches
        for path in (N.PATH_STREAM, N.PATH_TENSOR, N.PATH_EXACT):
            check_hybrid(idx.hybrid(Q, rb.hybrid_opts(10, 8, 0.3, path=path), kw), range(B), f"batch-{B} path {path}")
            top = idx.query(Q, 23, path=path)
            for b in range(B):
                ei, es = oracle.topk(X, Q[b], 23)
                if not (np.array_equal(top.row(b)[0], ei) and np.array_equal(top.row(b)[1], es)):
                    bad(f"top-23 rows={total} path={path} query={b}")
        # forced escalation is collective: every rank must take the same decisions
        r = idx.query(Q[:8], 10, path=N.PATH_STREAM, epsilon=10.0)
        for b in range(8):
            ei, es = oracle.topk(X, Q[b], 10)
            if not (np.array_equal(r.row(b)[0], ei) and np.array_equal(r.row(b)[1], es) and r.certified[b]):
                bad(f"escalation rows={total} query={b}")
        # MemoryStore.retrieve over the shards
        mem = idx.memory_retrieve(Q[:6], 5, 0.3, now_ms=go.now_ms)
        ct, cf, ac, la = oracle.gen_meta(go, 0, total)
        for b in range(6):
            vi, vs = oracle.topk(X, Q[b], 10)
            sel = vi.astype(np.int64)
            oi, osc, _ = oracle.memory_rank(vs, ct[sel] == 1, cf[sel], ac[sel], la[sel], go.now_ms, 5, 0.3)
            if not (np.array_equal(mem["ids"][b, :len(oi)], vi[oi]) and int(mem["counts"][b]) == len(oi)):
                bad(f"memory_retrieve rows={total} query={b}")
        # the micro-batcher forms batches from one process's own arrivals: on a shard it would desynchronise the ranks — refused
        try:
            rb.Batcher(idx, rb.hybrid_opts(10, 8, 0.3)).close()
            bad("a batcher was created on a shard")
        except N.RagError as e:
            if e.code != N.ERR_UNSUPPORTED:
                bad(f"batcher on a shard: error {e.code}")
        # a batch beyond the exported capacity is refused, not silently truncated
        try:
            idx.query(np.tile(Q, (2, 1)), 10, path=N.PATH_STREAM)
            bad("oversized batch was accepted")
        except N.RagError as e:
            if e.code != N.ERR_STATE:
                bad(f"oversized batch: error {e.code}")
        leave(idx)
        idx.close()

    # ---- a peer that never arrives: timeout, not a hang ------------------------------------------------
    total, d = 20_000, 128
    go = oracle.make_gen(total, n_clusters=16)
    gn = N.GenDesc.from_buffer_copy(bytes(go))
    base, n = shard_range(total, world, rank)
    idx = rb.VectorIndex(d, n, device=0, id_base=base)
    idx.generate(gn, n)
    bootstrap(idx, 32, 16, 120_000)
    Q = idx.generate_queries(gn, 0, 4)
    X = oracle.gen_rows(go, 0, total, d)
    idx.query(Q[:1], 10, path=N.PATH_STREAM)               # one healthy exchange
    leave(idx)
    bootstrap(idx, 32, 16, 1500)                           # now with a short fuse
    idx.query(Q[:1], 10, path=N.PATH_STREAM)
    rv.barrier()
    if rank == 0:
        t0 = time.time()
        try:
            idx.query(Q[1:2], 10, path=N.PATH_STREAM)      # the peers skip this call
            bad("a lone rank's exchange did not fail")
        except N.RagError as e:
            if e.code != N.ERR_TIMEOUT:
                bad(f"lone exchange: error {e.code} ({e})")
        waited = time.time() - t0
        if not 0.5 < waited < 60:
            bad(f"timeout took {waited:.1f} s")
        try:
            idx.query(Q[1:2], 10, path=N.PATH_STREAM)
            bad("a broken communicator accepted a call")
        except N.RagError as e:
            if e.code != N.ERR_STATE:
                bad(f"broken communicator: error {e.code}")
    leave(idx)
    bootstrap(idx, 32, 16, 120_000)                        # fresh mailboxes and counters on every rank
    r = idx.query(Q, 10, path=N.PATH_STREAM)
    for b in range(4):
        ei, es = oracle.topk(X, Q[b], 10)
        if not (np.array_equal(r.row(b)[0], ei) and np.array_equal(r.row(b)[1], es)):
            bad(f"after re-bootstrap query={b}")
    leave(idx)
    idx.close()
    print(f"[rank {rank}] " + ("ranks-on-one-gpu parity OK" if ok else "ranks-on-one-gpu parity FAILED"), flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

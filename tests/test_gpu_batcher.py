"""§8f N4 — the micro-batcher: concurrent batch-1 callers share corpus passes and every caller gets exactly
the result of a direct call."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_batcher_results_equal_direct_calls(native, oracle):
    import rag_era_b200 as rb

    n, d, nq, nthreads = 40000, 512, 256, 32
    go = oracle.make_gen(n, n_clusters=64, dup_period=19)
    gn = native.GenDesc.from_buffer_copy(bytes(go))
    with rb.VectorIndex(d, n, bf16_shadow=True) as idx:
        idx.generate(gn, n)
        Q = idx.generate_queries(gn, 0, nq)
        rng = np.random.default_rng(3)
        kw = [rng.integers(0, n, int(rng.integers(0, 6))).tolist() for _ in range(nq)]
        o = rb.hybrid_opts(10, 5, 0.3)
        direct = [idx.hybrid(Q[i], o, [kw[i]]).row(0) for i in range(nq)]     # batch-1 calls, one by one
        results = [None] * nq
        with rb.Batcher(idx, o, max_batch=64, max_wait_us=2000) as bt:
            def work(t):
                for i in range(t, nq, nthreads):
                    results[i] = bt.submit(Q[i], kw[i])
            th = [threading.Thread(target=work, args=(t,)) for t in range(nthreads)]
            [x.start() for x in th]
            [x.join() for x in th]
            st = bt.stats()
        assert st["queries"] == nq and st["batches"] < nq and st["largest_batch"] > 1      # requests were grouped
        for i in range(nq):
            for key in ("keys", "scores", "source", "ctype", "vec_ids", "vec_scores", "used_rrf"):
                assert np.array_equal(results[i][key], direct[i][key]), (i, key)
            e = oracle.hybrid_search(oracle.gen_rows(go, 0, n, d), Q[i], 10, 0.3, kw[i]) if i < 4 else None
            if e is not None:
                assert np.array_equal(results[i]["keys"], e["keys"]) and np.array_equal(results[i]["scores"], e["scores"])


def test_batcher_errors(native):
    import rag_era_b200 as rb

    with rb.VectorIndex(64, 100) as idx:
        idx.upload(np.ones((10, 64), np.float32))
        with pytest.raises(rb.RagError):
            rb.Batcher(idx, rb.hybrid_opts(0, 5, 0.3))
        with rb.Batcher(idx, rb.hybrid_opts(5, 2, 0.0), max_batch=8, max_wait_us=100) as bt:
            with pytest.raises(rb.RagError):
                bt.submit(np.ones(64, np.float32), [1, 2, 3])          # more keyword hits than keyword_limit
            r = bt.submit(np.ones(64, np.float32), [1])
            assert len(r["vec_ids"]) == 5


def test_batcher_async_submits_from_one_thread(native, oracle):
    """rag_batcher_submit_async: ONE thread keeps every request in flight (what Node's event loop does — no pool thread
    parked per request); results arrive through the completion callback and equal the direct batch-1 calls; batches form;
    RAG_ERR_BUSY (every batch buffer in flight) queues nothing (provoked deterministically in tests/c/batcher_tsan.cc)."""
    import time

    import rag_era_b200 as rb

    n, d, nq = 40000, 512, 192
    go = oracle.make_gen(n, n_clusters=64, dup_period=19)
    gn = native.GenDesc.from_buffer_copy(bytes(go))
    with rb.VectorIndex(d, n, bf16_shadow=True) as idx:
        idx.generate(gn, n)
        Q = idx.generate_queries(gn, 0, nq)
        rng = np.random.default_rng(5)
        kw = [rng.integers(0, n, int(rng.integers(0, 6))).tolist() for _ in range(nq)]
        o = rb.hybrid_opts(10, 5, 0.3)
        direct = [idx.hybrid(Q[i], o, [kw[i]]).row(0) for i in range(nq)]
        results, busy = [None] * nq, 0
        done = threading.Semaphore(0)
        with rb.Batcher(idx, o, max_batch=16, max_wait_us=2000) as bt:       # 4 buffers x 16 slots < 192 requests: BUSY may happen
            for i in range(nq):
                def on_done(r, i=i):
                    results[i] = r
                    done.release()
                while not bt.submit_async(Q[i], kw[i], on_done):
                    busy += 1
                    time.sleep(0.0005)
            for _ in range(nq):
                assert done.acquire(timeout=60)
            st = bt.stats()
            with pytest.raises(rb.RagError):
                bt.submit_async(Q[0], list(range(9)), lambda r: None)          # more keyword hits than keyword_limit
        assert st["queries"] == nq and st["largest_batch"] > 1 and st["batches"] < nq
        for i in range(nq):
            assert not isinstance(results[i], Exception), results[i]
            for key in ("keys", "scores", "source", "ctype", "vec_ids", "vec_scores", "used_rrf"):
                assert np.array_equal(results[i][key], direct[i][key]), (i, key)


def test_batcher_serves_the_retriever_seam(native, oracle):
    """`index.asRetriever({similarityTopK}).retrieve(q)` (hybrid-search.ts:223-224, memory/store.ts:111-116, the summarize tool)
    is a plain top-k. Through the batcher it is the vector-only branch with no keyword list and a filter below every cosine
    (keywordLimit 0, minVectorScore -2): ids and raw cosines equal rag_search's — what NativeVectorStore.query sends when
    RAGERA_BATCH=1."""
    import rag_era_b200 as rb

    n, d, nq, k = 20000, 256, 48, 7
    go = oracle.make_gen(n, n_clusters=32)
    gn = native.GenDesc.from_buffer_copy(bytes(go))
    with rb.VectorIndex(d, n, shadow="f16") as idx:
        idx.generate(gn, n)
        Q = idx.generate_queries(gn, 0, nq)
        direct = idx.query(Q, k)
        results = [None] * nq
        done = threading.Semaphore(0)
        with rb.Batcher(idx, rb.hybrid_opts(k, 0, -2.0), max_batch=64, max_wait_us=2000) as bt:
            for i in range(nq):
                def on_done(r, i=i):
                    results[i] = r
                    done.release()
                assert bt.submit_async(Q[i], [], on_done)
            for _ in range(nq):
                assert done.acquire(timeout=60)
        X = oracle.gen_rows(go, 0, n, d)
        for i in range(nq):
            r = results[i]
            assert not isinstance(r, Exception), r
            assert not r["used_rrf"] and r["certified"]
            assert np.array_equal(r["vec_ids"], direct.row(i)[0]) and np.array_equal(r["vec_scores"], direct.row(i)[1]), i
            assert np.array_equal(r["keys"], direct.row(i)[0]) and np.array_equal(r["scores"], direct.row(i)[1]), i
            if i < 4:
                ei, es = oracle.topk(X, Q[i], k)
                assert np.array_equal(r["vec_ids"], ei) and np.array_equal(r["vec_scores"], es)

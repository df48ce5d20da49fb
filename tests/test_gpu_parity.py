"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bar: chunk ids, fused order, source flags and RRF scores identical; cosine scores
bit-identical (K4 recomputes them in the reference's fp64 order), freshness within 2 ulp
(exp/log come from different libms). Run on a B200: ``pytest -m gpu``.
"""
import importlib
import json
import os
import struct

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CT = {"document": 0, "memory": 1, "code": 2}


def ulp_diff(a: float, b: float) -> int:
    ia, ib = struct.unpack("<q", struct.pack("<d", a))[0], struct.unpack("<q", struct.pack("<d", b))[0]
    return abs(ia - ib)


@pytest.fixture(scope="module")
def rb(native):
    import rag_era_b200

    assert native.load().rag_device_count() > 0, "no GPU visible: the gpu suite needs a B200"
    return rag_era_b200


def gen(oracle, native, total, **kw):
    g = oracle.make_gen(total, **kw)
    return g, native.GenDesc.from_buffer_copy(bytes(g))


# ------------------------------------------------------------------------------------------
# synthetic data: device generator == host generator, bit for bit
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype_name,d", [("f32", 1536), ("bf16", 1536), ("f32", 100), ("bf16", 72)])
def test_generator_bit_exact(rb, native, oracle, dtype_name, d):
    dt = native.F32 if dtype_name == "f32" else native.BF16
    go, gn = gen(oracle, native, 4096, n_clusters=64, dup_period=9, memory_rows=50)
    with rb.VectorIndex(d, 700, dtype=dt, id_base=1000) as idx:
        idx.generate(gn, 700)
        X = idx.read_rows(0, 700)
        E = oracle.gen_rows(go, 1000, 700, d, dtype=oracle.F32 if dt == native.F32 else oracle.BF16)
        assert np.array_equal(X, E)
        q = idx.generate_queries(gn, 5, 9)
        assert np.array_equal(q, oracle.gen_queries(go, 5, 9, d))


# ------------------------------------------------------------------------------------------
# top-k: ids identical, scores bit-identical
# ------------------------------------------------------------------------------------------
def check_topk(rb, native, oracle, X, Q, k, path, id_base=0, dtype=None, slack=0):
    dt = native.F32 if X.dtype == np.float32 else native.BF16
    with rb.VectorIndex(X.shape[1], max(len(X), 1), dtype=dt, id_base=id_base) as idx:
        idx.upload(X)
        r = idx.query(Q, k, path=path, slack=slack)
        for b in range(len(Q)):
            ei, es = oracle.topk(X, Q[b], k, id_base=id_base)
            gi, gs = r.row(b)
            assert np.array_equal(gi, ei), (b, gi, ei)
            assert np.array_equal(gs.view(np.uint64), es.view(np.uint64)), (b, gs, es)
            assert r.certified[b] == 1
        return r


@pytest.mark.parametrize("B,k", [(40, 10), (70, 29)])
def test_stream_path_batches_above_32_go_through_the_k3_merge(rb, native, oracle, B, k):
    """B > 32 on the stream path: per-CTA lists -> K3 tournament merge -> K4 (throughput variant)."""
    rng = np.random.default_rng(B)
    X = rng.standard_normal((6000, 96)).astype(np.float32)
    X[100] = X[7]                                                # an exact tie across CTAs
    Q = rng.standard_normal((B, 96)).astype(np.float32)
    check_topk(rb, native, oracle, X, Q, k, native.PATH_STREAM)


@pytest.mark.parametrize("path", ["stream", "exact"])
@pytest.mark.parametrize("n,d,k", [(1, 64, 5), (3, 64, 10), (31, 128, 29), (33, 96, 64), (1000, 1536, 10),
                                   (5000, 1024, 5), (4097, 100, 12), (20000, 256, 23), (257, 1536, 2)])
def test_topk_matches_oracle(rb, native, oracle, path, n, d, k):
    rng = np.random.default_rng(n * 31 + d)
    X = rng.standard_normal((n, d)).astype(np.float32) * rng.uniform(0.2, 3.0, (n, 1)).astype(np.float32)
    Q = (X[rng.integers(0, n, 3)] + 0.4 * rng.standard_normal((3, d))).astype(np.float32)
    check_topk(rb, native, oracle, X, Q, k, native.PATH_STREAM if path == "stream" else native.PATH_EXACT, id_base=7)


@pytest.mark.parametrize("path", ["stream", "exact"])
def test_topk_bf16_corpus(rb, native, oracle, path):
    rng = np.random.default_rng(5)
    X = oracle.f32_to_bf16(rng.standard_normal((3000, 512)).astype(np.float32))
    Q = (oracle.bf16_to_f32(X[[5, 77, 2999]]) + 0.3 * rng.standard_normal((3, 512))).astype(np.float32)
    check_topk(rb, native, oracle, X, Q, 10, native.PATH_STREAM if path == "stream" else native.PATH_EXACT)


@pytest.mark.parametrize("path", ["stream", "exact"])
def test_exact_ties_go_to_the_lower_id(rb, native, oracle, path):
    """Duplicate rows score identically in fp64; the reference's stable sort keeps insertion order."""
    go, gn = gen(oracle, native, 6000, n_clusters=4, dup_period=3)
    X = oracle.gen_rows(go, 0, 6000, 128)
    Q = oracle.gen_queries(go, 0, 6, 128)
    r = check_topk(rb, native, oracle, X, Q, 10, native.PATH_STREAM if path == "stream" else native.PATH_EXACT)
    ids, sc = r.row(0)
    ties = [(int(ids[i]), int(ids[i + 1])) for i in range(len(ids) - 1) if sc[i] == sc[i + 1]]
    assert ties and all(a < b for a, b in ties)


def test_many_identical_rows_escalate_and_stay_exact(rb, native, oracle):
    """More exact ties than the candidate window can hold: the stream path cannot certify, the
    library escalates (stream → exact → exact with K'=128) and the answer is still the oracle's."""
    rng = np.random.default_rng(9)
    X = rng.standard_normal((500, 64)).astype(np.float32)
    X[100:160] = X[100]                       # 60 identical rows
    Q = (X[100] + 0.05 * rng.standard_normal(64)).astype(np.float32)[None, :]
    check_topk(rb, native, oracle, X, Q, 10, native.PATH_STREAM)
    with rb.VectorIndex(64, 500) as idx:      # without escalation the query is flagged, not silently wrong
        idx.upload(X)
        r = idx.query(Q, 10, path=native.PATH_STREAM, flags=native.SEARCH_NO_ESCALATE)
        assert r.certified[0] == 0


def test_forced_escalation_matches(rb, native, oracle):
    rng = np.random.default_rng(21)
    X = rng.standard_normal((3000, 256)).astype(np.float32)
    Q = rng.standard_normal((5, 256)).astype(np.float32)
    with rb.VectorIndex(256, 3000) as idx:
        idx.upload(X)
        r = idx.query(Q, 8, path=native.PATH_STREAM, epsilon=10.0)      # nothing can be certified at eps=10
        for b in range(5):
            ei, es = oracle.topk(X, Q[b], 8)
            assert np.array_equal(r.row(b)[0], ei) and np.array_equal(r.row(b)[1], es)
        assert r.certified.all()                                        # re-run on the exact path


def test_append_rows_like_index_insert(rb, native, oracle):
    """index.insert appends one row at a time (memory/store.ts:67); order = insertion order."""
    rng = np.random.default_rng(2)
    X = rng.standard_normal((300, 64)).astype(np.float32)
    q = rng.standard_normal((1, 64)).astype(np.float32)
    with rb.VectorIndex(64, 400) as idx:
        idx.upload(X[:250])
        for i in range(250, 300):
            assert idx.upload(X[i:i + 1]) == i
        assert idx.rows == 300
        ei, es = oracle.topk(X, q[0], 7)
        gi, gs = idx.query(q, 7).row(0)
        assert np.array_equal(gi, ei) and np.array_equal(gs, es)


def test_api_errors(rb, native):
    with rb.VectorIndex(64, 10) as idx:
        with pytest.raises(rb.RagError) as e:
            idx.query(np.zeros((1, 64), np.float32), 5)
        assert e.value.code == native.ERR_STATE                 # empty index
        idx.upload(np.ones((4, 64), np.float32))
        with pytest.raises(rb.RagError) as e:
            idx.query(np.zeros((1, 64), np.float32), 65)
        assert e.value.code == native.ERR_INVALID
        r = idx.query(np.ones((1, 64), np.float32), 4, path=native.PATH_TENSOR)   # fp32 index, no shadow: tf32 path
        assert r.counts[0] == 4
        with pytest.raises(rb.RagError):
            idx.upload(np.ones((20, 64), np.float32))           # exceeds capacity


# ------------------------------------------------------------------------------------------
# RRF / filter / hybrid
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("path", ["stream", "exact", "tensor"])
def test_topk_kernels_on_golden_vectors(rb, native, golden, path):
    """K1 / K1x / K2 (+K3, K4) against the golden cosine/top-k cases of tests/golden/make_kat.py: ids in the reference's
    order (exact ties to the earlier row) and every similarity bit for bit."""
    p = {"stream": native.PATH_STREAM, "exact": native.PATH_EXACT, "tensor": native.PATH_TENSOR}[path]
    for c in golden["cosine_topk"]:
        X = np.array([[float.fromhex(v) for v in r] for r in c["rows"]], dtype=np.float32)
        q = np.array([float.fromhex(v) for v in c["query"]], dtype=np.float32)
        with rb.VectorIndex(c["dim"], len(X), bf16_shadow=(path == "tensor")) as idx:
            idx.upload(X)
            r = idx.query(q, c["k"], path=p)
            ids, sc = r.row(0)
            assert [int(i) for i in ids] == c["ids"], (path, c["name"])
            assert [float(s).hex() for s in sc] == c["similarities"], (path, c["name"])
            assert r.certified[0], (path, c["name"])


def test_rrf_kernel_on_golden_vectors(rb, native, golden):
    SRC = {0: "vector", 1: "keyword", 2: "both"}
    with rb.VectorIndex(64, 4) as idx:
        for case in golden["rrf"]:
            c = case["config"]
            cfg = rb.RRFConfig(c["k"], c["vectorWeight"], c["keywordWeight"], c["bothBonus"])
            vk = [v[0] for v in case["vector"]]
            vt = [CT[v[1]] for v in case["vector"]]
            r = idx.rrf_fuse([vk], [case["keyword"]], cfg, vec_ctypes=[vt]).row(0)
            exp = case["expect"]
            assert len(r["keys"]) == len(exp), case["name"]
            for i, e in enumerate(exp):
                assert int(r["keys"][i]) == e["key"], (case["name"], i)
                assert float(r["scores"][i]).hex() == float.fromhex(e["hex"]).hex(), (case["name"], i)
                assert SRC[int(r["source"][i])] == e["source"] and int(r["ctype"][i]) == CT[e["contentType"]]
        # batched: all cases in one launch give the same answers
        cfg = rb.RRFConfig()
        doc_cases = [c for c in golden["rrf"] if c["config"] == dict(k=60, vectorWeight=1.0, keywordWeight=1.0, bothBonus=0.1)]
        out = idx.rrf_fuse([[v[0] for v in c["vector"]] for c in doc_cases], [c["keyword"] for c in doc_cases], cfg)
        for b, c in enumerate(doc_cases):
            assert [int(x) for x in out.row(b)["keys"]] == [e["key"] for e in c["expect"]]
            assert [float(x).hex() for x in out.row(b)["scores"]] == [float.fromhex(e["hex"]).hex() for e in c["expect"]]


@pytest.mark.parametrize("dtype_name", ["f32", "bf16"])
def test_hybrid_matches_oracle(rb, native, oracle, dtype_name):
    dt = native.F32 if dtype_name == "f32" else native.BF16
    n, d = 8000, 1536
    go, gn = gen(oracle, native, n, n_clusters=32, dup_period=11, memory_rows=800)
    X = oracle.gen_rows(go, 0, n, d, dtype=oracle.F32 if dt == native.F32 else oracle.BF16)
    B = 12
    Q = oracle.gen_queries(go, 0, B, d)
    rng = np.random.default_rng(17)
    row_keys = np.arange(n, dtype=np.uint64)
    row_keys[rng.integers(0, n, 2000)] = rng.integers(0, 50, 2000).astype(np.uint64)   # key collisions (duplicate content prefixes)
    ctype = np.where(np.arange(n) < 800, 1, np.where(np.arange(n) % 5 == 0, 2, 0)).astype(np.uint8)
    with rb.VectorIndex(d, n, dtype=dt) as idx:
        idx.upload(X)
        idx.set_row_meta(0, content_type=ctype, confidence=np.ones(n), access_count=np.zeros(n, np.int32),
                         last_access_ms=np.zeros(n, np.int64))
        idx.set_row_keys(0, row_keys)
        for (vk, kl, thr, cfg) in [(5, 5, 0.3, rb.RRFConfig()), (10, 10, 0.3, rb.RRFConfig()),
                                   (23, 8, 0.4, rb.RRFConfig()), (6, 5, 0.25, rb.RRFConfig(40, 1.0, 1.3, 0.15)),
                                   (8, 0, 0.3, rb.RRFConfig()), (29, 19, 0.95, rb.RRFConfig())]:
            kw_lists = []
            for b in range(B):
                top_i, _ = oracle.topk(X, Q[b], vk)
                hits = [int(row_keys[i]) for i in top_i[: max(1, (3 * kl + 9) // 10)]] if kl else []
                hits += rng.integers(0, n, max(0, kl - len(hits))).tolist()
                rng.shuffle(hits)
                if b % 4 == 3:
                    hits = []                                   # Meilisearch down / no hits → vector-only branch
                kw_lists.append(hits[:kl])
            o = rb.hybrid_opts(vk, kl, thr, cfg, path=native.PATH_STREAM)
            res = idx.hybrid(Q, o, kw_lists)
            ocfg = oracle.RRFConfig(cfg.k, cfg.vector_weight, cfg.keyword_weight, cfg.both_bonus)
            for b in range(B):
                e = oracle.hybrid_search(X, Q[b], vk, thr, kw_lists[b], ocfg, row_keys=row_keys, row_ctype=ctype)
                g = res.row(b)
                assert g["used_rrf"] == e["used_rrf"], (vk, kl, b)
                assert np.array_equal(g["vec_ids"], e["vec_ids"]) and np.array_equal(g["vec_scores"], e["vec_scores"])
                assert np.array_equal(g["keys"], e["keys"]), (vk, kl, b, g["keys"], e["keys"])
                assert np.array_equal(g["scores"].view(np.uint64), e["scores"].view(np.uint64)), (vk, kl, b)
                assert np.array_equal(g["source"], e["source"]) and np.array_equal(g["ctype"], e["ctype"])
                assert g["certified"]


def test_staged_form_equals_direct_call(rb, native, oracle):
    go, gn = gen(oracle, native, 5000, n_clusters=16)
    with rb.VectorIndex(256, 5000) as idx:
        idx.generate(gn, 5000)
        Q = idx.generate_queries(gn, 0, 4)
        kw = [[1, 2, 3], [], [4999, 0], [7]]
        o = rb.hybrid_opts(10, 10, 0.3, path=native.PATH_STREAM)
        a = idx.hybrid(Q, o, kw)
        B = idx.stage_batch(Q, kw, 10)
        idx.timer_start()
        n0 = idx.launch_count
        idx.hybrid_staged(B, o)
        ms = idx.timer_stop()
        assert ms > 0 and idx.launch_count - n0 == 2          # K1, then K3+K4+K5 in one kernel (small batch, one GPU)
        b = idx.fetch_fused(B, o)
        for q in range(4):
            for key in ("keys", "scores", "source", "ctype", "vec_ids", "vec_scores"):
                assert np.array_equal(a.row(q)[key], b.row(q)[key])


def test_third_list_extension_reduces_to_reference(rb, native, oracle):
    """SURVEY N-c4: with fresh_limit=0 the result is the reference's two-list RRF; with a
    freshness list it equals the oracle's rrf3 over memory hits ranked by freshness."""
    n, d, now = 6000, 128, 1_760_000_000_000
    go, gn = gen(oracle, native, n, n_clusters=8, memory_rows=3000, now_ms=now)
    X = oracle.gen_rows(go, 0, n, d)
    ct, cf, ac, la = oracle.gen_meta(go, 0, n)
    Q = oracle.gen_queries(go, 0, 6, d)
    with rb.VectorIndex(d, n) as idx:
        idx.generate(gn, n)
        kw = [[int(x) for x in oracle.topk(X, Q[b], 3)[0]] + [5999, 5998] for b in range(6)]
        o3 = rb.hybrid_opts(12, 5, 0.2, path=native.PATH_STREAM, fresh_limit=6, fresh_weight=0.8, now_ms=now)
        r3 = idx.hybrid(Q, o3, kw)
        for b in range(6):
            vi, vs = oracle.topk(X, Q[b], 12)
            vi, vs = oracle.filter_min_score(vi, vs, 0.2)
            mem = [int(i) for i in vi if ct[int(i)] == 1]
            fr = {i: oracle.freshness(cf[i], int(ac[i]), int(la[i]), now) for i in mem}
            fl = sorted(mem, key=lambda i: (-fr[i], i))[:6]
            ek, es, esrc, ect = oracle.rrf(vi, kw[b], oracle.RRFConfig(), vec_ctype=[ct[int(i)] for i in vi],
                                           fresh_keys=fl, fresh_weight=0.8)
            g = r3.row(b)
            assert np.array_equal(g["keys"], ek) and np.array_equal(g["scores"], es)
            assert np.array_equal(g["source"], esrc)


# ------------------------------------------------------------------------------------------
# memory: freshness + MemoryStore.retrieve
# ------------------------------------------------------------------------------------------
def test_freshness_kernel(rb, native, oracle, golden):
    now = 1_760_000_000_000
    rng = np.random.default_rng(4)
    n = 5000
    conf = rng.uniform(0.0, 1.0, n)
    acc = rng.integers(0, 200, n).astype(np.int32)
    last = now - rng.integers(0, 400 * 3600000, n)
    with rb.VectorIndex(64, 4) as idx:
        got = idx.freshness_scores(conf, acc, last, now)
        exp = np.array([oracle.freshness(conf[i], int(acc[i]), int(last[i]), now) for i in range(n)])
        assert max(ulp_diff(float(a), float(b)) for a, b in zip(got, exp)) <= 2     # exp/log: libm variance
        assert ((got >= 0) & (got <= 1)).all()
        for e in golden["freshness"]:
            v = idx.freshness_scores([e["confidence"]], [e["accessCount"]], [e["lastAccessedMs"]], e["nowMs"])[0]
            assert ulp_diff(float(v), float.fromhex(e["hex"])) <= 2


def test_memory_retrieve_matches_oracle(rb, native, oracle):
    n, d, now = 4000, 256, 1_760_000_000_000
    go, gn = gen(oracle, native, n, n_clusters=4, memory_rows=1500, now_ms=now, query_noise=0.2)
    X = oracle.gen_rows(go, 0, n, d)
    ct, cf, ac, la = oracle.gen_meta(go, 0, n)
    Q = oracle.gen_queries(go, 0, 8, d)
    with rb.VectorIndex(d, n) as idx:
        idx.generate(gn, n)
        for limit, minrel in [(10, 0.5), (1, 0.5), (5, 0.3), (32, 0.0)]:
            r = idx.memory_retrieve(Q, limit, minrel, now_ms=now, path=native.PATH_STREAM)
            for b in range(8):
                vi, vs = oracle.topk(X, Q[b], 2 * limit)
                oi, osc, ofr = oracle.memory_rank(vs, [ct[int(i)] for i in vi], [cf[int(i)] for i in vi],
                                                  [ac[int(i)] for i in vi], [la[int(i)] for i in vi], now, limit, minrel)
                cnt = int(r["counts"][b])
                assert cnt == len(oi)
                assert np.array_equal(r["ids"][b, :cnt], vi[oi])
                assert np.array_equal(r["relevance"][b, :cnt], vs[oi])
                assert max([ulp_diff(float(a), float(e)) for a, e in zip(r["freshness"][b, :cnt], ofr)] + [0]) <= 2
                assert np.allclose(r["scores"][b, :cnt], osc, rtol=0, atol=1e-15)


# ------------------------------------------------------------------------------------------
# the reference-shaped host API end to end (strings in, HybridSearchResult out)
# ------------------------------------------------------------------------------------------
def test_hybrid_search_host_mirror(rb, native, oracle):
    hs = importlib.import_module("rag_era_b200.hybrid_search")
    rng = np.random.default_rng(33)
    d, n = 128, 400
    E = rng.standard_normal((n, d)).astype(np.float32)
    nodes = [hs.Node(f"node-{i}", f"【文档: doc{i % 7}.md】\n\nchunk {i} " + "x" * (i % 5) * 30,
                     dict(documentName=f"doc{i % 7}.md")) for i in range(n)]
    nodes[10].metadata = dict(type="memory", memoryId="m1")
    nodes[11].metadata = dict(relativePath="src/a.ts", language="ts")
    vocab = {"alpha": E[10] + 0.05 * rng.standard_normal(d).astype(np.float32)}
    index = hs.KnowledgeIndex(d, n, embed_model=lambda s: vocab[s])
    try:
        index.insert_nodes(nodes, E)

        class Meili:
            up = True

            def is_available(self):
                return self.up

            def search(self, kb, query, limit):
                hits = [hs.KeywordHit("h0", "D3", "doc3.md", nodes[10].text), hs.KeywordHit("h1", "D9", "other.md", "**alpha** raw text"),
                        hs.KeywordHit("h2", "D4", "doc4.md", nodes[11].text)]
                return hits[:limit]

        m = Meili()
        res = hs.hybrid_search(index, "kb1", "alpha", dict(vectorTopK=5, keywordLimit=5, minVectorScore=0.0), keyword_service=m)
        ei, es = oracle.topk(E, vocab["alpha"], 5)
        keys = hs.KeyInterner()
        vk = [keys.key(nodes[int(i)].text) for i in ei]
        kk = [keys.key(t) for t in (nodes[10].text, "**alpha** raw text", nodes[11].text)]
        ek, esc, esrc, _ = oracle.rrf(vk, kk)
        assert [r.id for r in res] == [keys.string(int(k)) for k in ek]
        assert [r.score for r in res] == list(esc)
        assert [r.source for r in res] == [hs.SOURCE[int(s)] for s in esrc]
        assert res[0].content == nodes[10].text and res[0].source == "both" and res[0].contentType == "memory"
        assert res[0].documentName == "用户记忆" and res[0].documentId is None
        kw_only = [r for r in res if r.source == "keyword"]
        assert any(r.documentId == "D9" and r.contentType == "document" for r in kw_only)
        # index.asRetriever({similarityTopK}).retrieve(query) — the seam every caller goes through (:223-224)
        hits = index.as_retriever(similarity_top_k=5).retrieve("alpha")
        assert [h.node.id_ for h in hits] == [f"node-{int(i)}" for i in ei] and [h.score for h in hits] == list(es)
        assert len(index.as_retriever().retrieve("alpha")) == 2                  # llamaindex default similarityTopK
        # Meilisearch down → vector-only branch: raw cosines, node ids
        m.up = False
        res2 = hs.hybrid_search(index, "kb1", "alpha", dict(vectorTopK=5, minVectorScore=0.0), keyword_service=m)
        assert [r.id for r in res2] == [f"node-{int(i)}" for i in ei] and [r.score for r in res2] == list(es)
        assert all(r.source == "vector" for r in res2)
        # keywordLimit 0 is respected (?? semantics) → vector-only as well
        m.up = True
        res3 = hs.hybrid_search(index, "kb1", "alpha", dict(vectorTopK=5, keywordLimit=0, minVectorScore=0.0), keyword_service=m)
        assert [r.id for r in res3] == [r.id for r in res2]
        # preset 'code' (hybrid-search.ts:98-104,296): RRF {40, 1, 1.3, 0.15}, vectorTopK 6, keywordLimit 5, min 0.25 — and
        # every non-memory VECTOR hit reads as 'code'; keyword-only hits stay 'document' (:178-187), memories stay memories
        res4 = hs.hybrid_search(index, "kb1", "alpha", dict(preset="code", minVectorScore=0.0), keyword_service=m)
        ei6, _ = oracle.topk(E, vocab["alpha"], 6)
        keys4 = hs.KeyInterner()
        vk4 = [keys4.key(nodes[int(i)].text) for i in ei6]
        kk4 = [keys4.key(t) for t in (nodes[10].text, "**alpha** raw text", nodes[11].text)]
        ek4, esc4, esrc4, _ = oracle.rrf(vk4, kk4, oracle.RRFConfig(40.0, 1.0, 1.3, 0.15))
        assert [r.id for r in res4] == [keys4.string(int(k)) for k in ek4] and [r.score for r in res4] == list(esc4)
        vec_texts = {nodes[int(i)].text for i in ei6}
        for r in res4:
            want = "memory" if r.content == nodes[10].text else "code" if r.content in vec_texts else "document"
            assert r.contentType == want, (r.content[:30], r.contentType, want)
        res5 = hs.hybrid_search(index, "codebase_7", "alpha", dict(vectorTopK=5, minVectorScore=0.0), keyword_service=m)
        assert [r.score for r in res5] == [r.score for r in res] and {r.contentType for r in res5} == {"memory", "code", "document"}
        # reciprocal_rank_fusion drop-in on result objects
        fused = hs.reciprocal_rank_fusion(res2, m.search("kb1", "alpha", 3), store=index.store)
        assert [r.score for r in fused] == [r.score for r in res]
    finally:
        index.close()


def test_context_engine_and_search_tools_mirror(rb, native, oracle):
    """The callers named by north_star — getUnifiedResults (engine.ts:225-299), search_knowledge / deep_search
    (search-tools.ts:12-95) — against the oracle's top-k + RRF on the same inputs."""
    hs = importlib.import_module("rag_era_b200.hybrid_search")
    rng = np.random.default_rng(41)
    d, n = 96, 600
    centre = rng.standard_normal(d).astype(np.float32)
    E = (centre + 0.9 * rng.standard_normal((n, d))).astype(np.float32)          # cosines to the query ~0.7: clear the 0.4 filter
    nodes = [hs.Node(f"node-{i}", f"retrieval augmented generation chunk {i}: " + " ".join(f"w{(i * 7 + j) % 97}" for j in range(12)),
                     dict(documentName=f"doc{i % 9}.md")) for i in range(n)]
    for i in range(0, n, 6):
        nodes[i].metadata = dict(type="memory", memoryId=f"mem-{i}", memoryType="preference")
    index = hs.KnowledgeIndex(d, n, embed_model=lambda s: centre)
    try:
        index.insert_nodes(nodes, E)
        vi29, _ = oracle.topk(E, centre, 29)

        class Meili:
            def is_available(self):
                return True

            def search(self, kb, query, limit):
                hits = [hs.KeywordHit(f"h{j}", f"D{j}", f"doc{j}.md", nodes[int(vi29[2 * j + 1])].text) for j in range(4)]
                hits.insert(2, hs.KeywordHit("hx", "DX", "other.md", "retrieval generation keyword-only hit with enough text"))
                return hits[:limit]

        def expect(top_k, kw_limit, min_score):
            vi, vs = oracle.topk(E, centre, top_k)
            keep = [(int(i), s) for i, s in zip(vi, vs) if s >= min_score]
            keys = hs.KeyInterner()
            vk = [keys.key(nodes[i].text) for i, _ in keep]
            hits = Meili().search("kb", "q", kw_limit)
            ek, esc, esrc, _ = oracle.rrf(vk, [keys.key(h.content) for h in hits])
            return keys, keep, hits, [int(k) for k in ek], list(esc), [int(s) for s in esrc]

        # --- ContextEngine.getUnifiedResults: default decision → hybridSearch(vectorTopK 18, keywordLimit 6, min 0.4)
        got = rb.get_unified_results(index, "kb1", "retrieval generation", keyword_service=Meili(), now_ms=123)
        keys, keep, hits, ek, esc, esrc = expect(18, 6, 0.4)
        by_key = {}
        for i, _ in keep:
            by_key.setdefault(keys.key(nodes[i].text), nodes[i])
        e_mem = [(k, s) for k, s in zip(ek, esc) if k in by_key and (by_key[k].metadata or {}).get("type") == "memory"]
        e_doc = [(k, s, src) for k, s, src in zip(ek, esc, esrc) if not (k in by_key and (by_key[k].metadata or {}).get("type") == "memory")]
        assert len(e_mem) >= 1 and len(e_doc) >= 5
        assert [(m.id, m.score) for m in got["memories"]] == [(by_key[k].metadata["memoryId"], s) for k, s in e_mem][:10]
        assert all(m.relevanceScore == m.score and m.freshnessScore == 0.5 and m.confidence == 0.8 and m.type == "preference"
                   for m in got["memories"])
        assert [(r.id, r.score) for r in got["raw_documents"]] == [(keys.string(k), s) for k, s, _ in e_doc]
        assert [r.source for r in got["raw_documents"]] == [{0: "vector", 1: "keyword", 2: "hybrid"}[s] for _, _, s in e_doc]
        assert got["documents"] == rb.process_results(got["raw_documents"], "retrieval generation") and len(got["documents"]) <= 10
        # a semantic high-priority decision asks for k = 19 + 10 = 29 and no keyword list → vector-only branch, raw cosines
        sem = rb.get_unified_results(index, "kb1", "retrieval generation", rb.RetrievalDecision(queryType="semantic", priority="high"),
                                     keyword_service=Meili())
        vi, vs = oracle.topk(E, centre, 29)
        e = [(int(i), s) for i, s in zip(vi, vs) if s >= 0.4]
        assert [r.score for r in sem["raw_documents"]] == [s for i, s in e if i % 6 != 0]
        assert [m.score for m in sem["memories"]] == [s for i, s in e if i % 6 == 0][:10]

        # --- the agent's tools: top-5 → show 3, top-10 → show 8 (preset minVectorScore 0.3)
        for tool, name, top_k, show in ((rb.search_knowledge, "search_knowledge", 5, 3), (rb.deep_search, "deep_search", 10, 8)):
            ctx = rb.ToolContext(index, "kb1", keyword_service=Meili())
            text = tool(ctx, "retrieval generation")
            keys, keep, hits, ek, esc, esrc = expect(top_k, top_k, 0.3)
            assert [r.id for r in ctx.searchResults] == [keys.string(k) for k in ek]
            assert [r.score for r in ctx.searchResults] == esc
            assert text == hs.format_search_results(ctx.searchResults, show)
            assert text.count("[来源") == min(show, len(ek)) and ctx.toolCalls[0]["tool"] == name
            assert ctx.toolCalls[0]["output"] == hs.js_substring(text, 0, 200)
            first = len(ctx.searchResults)
            tool(ctx, "retrieval generation")                                    # results are saved once (:33-35)
            assert len(ctx.searchResults) == first and len(ctx.toolCalls) == 2
    finally:
        index.close()


def test_memory_store_mirror(rb, native, oracle):
    hs = importlib.import_module("rag_era_b200.hybrid_search")
    rng = np.random.default_rng(8)
    d, now = 64, 1_760_000_000_000
    E = rng.standard_normal((50, d)).astype(np.float32)
    index = hs.KnowledgeIndex(d, 100)
    try:
        index.insert_nodes([hs.Node(f"n{i}", f"doc chunk {i}", dict(documentName="d.md")) for i in range(50)], E)
        ms = rb.MemoryStore("kb1", index)
        mems = [rb.Memory(f"m{i}", "kb1", f"user likes {i}", 0.5 + 0.05 * i, i, now - i * 7_200_000) for i in range(6)]
        embs = [(E[3] + 0.1 * (i + 1) * rng.standard_normal(d)).astype(np.float32) for i in range(6)]
        for m, e in zip(mems, embs):
            ms.store(m, e)
        q = E[3]
        got = ms.retrieve(q, 3, 0.5, now_ms=now)
        allX = np.vstack([E] + [e[None, :] for e in embs])
        vi, vs = oracle.topk(allX, q, 6)
        ism = [1 if i >= 50 else 0 for i in vi]
        conf = [mems[int(i) - 50].confidence if i >= 50 else 0 for i in vi]
        acc = [mems[int(i) - 50].accessCount if i >= 50 else 0 for i in vi]
        la = [mems[int(i) - 50].lastAccessedAt if i >= 50 else 0 for i in vi]
        oi, osc, ofr = oracle.memory_rank(vs, ism, conf, acc, la, now, 3, 0.5)
        assert [g.id for g in got] == [mems[int(vi[i]) - 50].id for i in oi]
        assert np.allclose([g.score for g in got], osc, rtol=0, atol=1e-15)
        assert [g.relevanceScore for g in got] == [vs[i] for i in oi]
        assert ms.has_similar(embs[0], 0.9, now) == (oracle.cosine(embs[0], embs[0]) >= 0.9)
        # sortByFreshness (freshness.ts:74-83): device freshness, stable order for ties (two identical memories keep their order)
        twins = mems + [rb.Memory("twin", "kb1", "same as m1", mems[1].confidence, mems[1].accessCount, mems[1].lastAccessedAt)]
        of = [oracle.freshness(m.confidence, m.accessCount, m.lastAccessedAt, now) for m in twins]
        want_order = [twins[i].id for i in sorted(range(len(twins)), key=lambda i: -of[i])]
        assert [m.id for m in rb.sort_by_freshness(twins, now, store=index.store)] == want_order
        assert want_order.index("m1") < want_order.index("twin")
        # touch (store.ts:207-215): accessCount + 1, lastAccessedAt = now — the device copy follows, freshness changes
        later = now + 3_600_000
        ms.touch_many([mems[2].id, mems[4].id], later)
        assert (mems[2].accessCount, mems[2].lastAccessedAt) == (3, later)
        acc2 = [mems[int(i) - 50].accessCount if i >= 50 else 0 for i in vi]
        la2 = [mems[int(i) - 50].lastAccessedAt if i >= 50 else 0 for i in vi]
        oi2, osc2, ofr2 = oracle.memory_rank(vs, ism, conf, acc2, la2, later, 3, 0.5)
        got2 = ms.retrieve(q, 3, 0.5, now_ms=later)
        assert [g.id for g in got2] == [mems[int(vi[i]) - 50].id for i in oi2]
        assert np.allclose([g.score for g in got2], osc2, rtol=0, atol=1e-15) and np.allclose([g.freshnessScore for g in got2], ofr2, rtol=0, atol=1e-15)
        # delete (store.ts:240-251) removes only the DB record: the vector node stays, retrieve skips it (`if (dbMemory)`, :153)
        # BEFORE the slice, so the next-best memory moves up instead of leaving a hole
        gone = got2[0].id
        ms.delete(gone)
        oi6, _, _ = oracle.memory_rank(vs, ism, conf, acc2, la2, later, 6, 0.5)
        want = [mems[int(vi[i]) - 50].id for i in oi6 if mems[int(vi[i]) - 50].id != gone][:3]
        got3 = ms.retrieve(q, 3, 0.5, now_ms=later)
        assert [g.id for g in got3] == want and len(want) == 3 and ms.count() == 5
        hs_res = hs.hybrid_search(index, "kb1", q, dict(vectorTopK=6, minVectorScore=0.0))
        assert any(r.contentType == "memory" and r.id == f"memory_{gone}" for r in hs_res)      # still a memory hit of hybridSearch
    finally:
        index.close()


# ------------------------------------------------------------------------------------------
# BASELINE sizes: size-independent properties + one exhaustive oracle comparison
# ------------------------------------------------------------------------------------------
def test_c2_full_size_properties(rb, native, oracle):
    """C2: 1M x 1536 fp32, deep_search (vectorTopK=10, keywordLimit=10, min 0.3, RRF k=60)."""
    n, d, B = 1_000_000, 1536, 16
    go, gn = gen(oracle, native, n)
    with rb.VectorIndex(d, n) as idx:
        idx.generate(gn, n)
        Q = idx.generate_queries(gn, 0, B)
        r = idx.query(Q, 10, path=native.PATH_STREAM)
        assert r.certified.all() and (r.counts == 10).all()
        planted = [int(oracle_planted(go, b)) for b in range(B)]
        for b in range(B):
            ids, sc = r.row(b)
            assert (np.diff(sc) <= 0).all()                                  # sorted
            assert int(ids[0]) == planted[b]                                 # the planted row wins
            rows = idx.read_rows(int(ids.min()), 1)                          # spot check: read back == generator
            assert np.array_equal(rows, oracle.gen_rows(go, int(ids.min()), 1, d))
            for i, s in zip(ids, sc):                                        # every reported score is the oracle's, bit for bit
                assert oracle.cosine(Q[b], oracle.gen_rows(go, int(i), 1, d)[0]) == s
        r2 = idx.query(Q, 10, path=native.PATH_STREAM)                       # idempotent
        assert np.array_equal(r.ids, r2.ids) and np.array_equal(r.scores, r2.scores)
        # one exhaustive comparison against the oracle scanning all 1M generated rows
        ei, es = oracle.topk_generated(go, oracle.F32, 0, n, d, Q[0], 10)
        assert np.array_equal(r.row(0)[0], ei) and np.array_equal(r.row(0)[1], es)
        # top-10 is a prefix-consistent family: top-5 == first 5 of top-10
        r5 = idx.query(Q, 5, path=native.PATH_STREAM)
        assert np.array_equal(r5.ids, r.ids[:, :5])
        # hybrid on top: fused order is sorted, 'both' keys are exactly the intersection
        kw = [[int(x) for x in r.ids[b, :3]] + [(planted[b] * 7 + j) % n for j in range(7)] for b in range(B)]
        f = idx.hybrid(Q, rb.hybrid_opts(10, 10, 0.3, path=native.PATH_STREAM), kw)
        for b in range(B):
            g = f.row(b)
            assert g["used_rrf"] and (np.diff(g["scores"]) <= 0).all()
            both = {int(k) for k, s in zip(g["keys"], g["source"]) if s == native.SRC_BOTH}
            assert both == set(int(x) for x in g["vec_ids"]) & set(kw[b])
            ek, es2, esrc, _ = oracle.rrf(g["vec_ids"], kw[b])
            assert np.array_equal(g["keys"], ek) and np.array_equal(g["scores"], es2)


def oracle_planted(g, b):
    """rg_planted_row of include/ragera_gen.h restated with Python integers."""
    M = (1 << 64) - 1

    def mix(z):
        z = (z + 0x9E3779B97F4A7C15) & M
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        return z ^ (z >> 31)

    return mix(g.query_seed ^ 0x9A17 ^ mix(b)) % g.total_rows


def exhaustive_topk(oracle, idx, go, Q, k, slab=400_000):
    """The oracle's getTopKEmbeddings over EVERY row the index holds, for each query of Q: the rows are read back
    from the device slab by slab (so the oracle scores exactly the bytes the kernels scored; a few rows of every
    slab are also checked against the host generator), scored with the reference's fp64 left-to-right chain on all
    host threads, and the per-slab top-k lists are merged on (score desc, chunk id asc)."""
    n, d = idx.rows, idx.dim
    odt = oracle.F32 if idx._np_dtype() == np.float32 else oracle.BF16
    ids = [[] for _ in range(len(Q))]
    scs = [[] for _ in range(len(Q))]
    rng = np.random.default_rng(n)
    for lo in range(0, n, slab):
        m = min(slab, n - lo)
        X = idx.read_rows(lo, m)
        for r in rng.integers(0, m, 3):
            assert np.array_equal(X[r], oracle.gen_rows(go, idx.id_base + lo + int(r), 1, d, dtype=odt)[0])
        for b in range(len(Q)):
            i, s = oracle.topk(X, Q[b], k, id_base=idx.id_base + lo)
            ids[b].append(i)
            scs[b].append(s)
        del X
    out = []
    for b in range(len(Q)):
        i, s = np.concatenate(ids[b]), np.concatenate(scs[b])
        order = np.lexsort((i, -s))[:k]
        out.append((i[order], s[order]))
    return out


@pytest.mark.parametrize("B", [2, 4, 5, 8, 9, 17])
@pytest.mark.parametrize("d", [1536, 100])
def test_multi_query_stream_kernel(rb, native, oracle, B, d):
    """K1m: several queries share one corpus pass; results must not depend on how queries are grouped."""
    n = 7000
    go, gn = gen(oracle, native, n, n_clusters=16, dup_period=6)
    X = oracle.gen_rows(go, 0, n, d)
    with rb.VectorIndex(d, n) as idx:
        idx.generate(gn, n)
        Q = idx.generate_queries(gn, 0, B)
        r = idx.query(Q, 10, path=native.PATH_STREAM)
        single = [idx.query(Q[b:b + 1], 10, path=native.PATH_STREAM) for b in range(B)]
        for b in range(B):
            ei, es = oracle.topk(X, Q[b], 10)
            assert np.array_equal(r.row(b)[0], ei) and np.array_equal(r.row(b)[1], es)
            assert np.array_equal(single[b].row(0)[0], ei)
        assert r.certified.all()


@pytest.mark.parametrize("B", [2, 4, 5, 9])
@pytest.mark.parametrize("d", [1536, 512, 72])
def test_multi_query_stream_kernel_bf16_corpus(rb, native, oracle, B, d):
    """K1m for a bf16 corpus (4 queries share one corpus pass): what escalated tensor-path queries of a bf16 index run."""
    n = 9000
    go, gn = gen(oracle, native, n, n_clusters=16, dup_period=6)
    X = oracle.gen_rows(go, 0, n, d, dtype=oracle.BF16)
    with rb.VectorIndex(d, n, dtype=native.BF16) as idx:
        idx.generate(gn, n)
        Q = idx.generate_queries(gn, 0, B)
        r = idx.query(Q, 10, path=native.PATH_STREAM)
        for b in range(B):
            ei, es = oracle.topk(X, Q[b], 10)
            assert np.array_equal(r.row(b)[0], ei) and np.array_equal(r.row(b)[1], es)
        assert r.certified.all()


def test_c3_full_size_properties(rb, native, oracle):
    """C3: 10M x 1536 fp32 (61 GB), batch 1 — the north-star target config: size-independent properties and an
    exhaustive oracle scan of the whole corpus for every query."""
    n, d, B = 10_000_000, 1536, 4
    go, gn = gen(oracle, native, n)
    with rb.VectorIndex(d, n) as idx:
        idx.generate(gn, n)
        Q = idx.generate_queries(gn, 0, B)
        r = [idx.query(Q[b:b + 1], 10, path=native.PATH_STREAM) for b in range(B)]     # batch 1, as benchmarked
        rb4 = idx.query(Q, 10, path=native.PATH_STREAM)                                # the multi-query kernel agrees
        for b in range(B):
            ids, sc = r[b].row(0)
            assert r[b].certified[0] == 1 and len(ids) == 10 and (np.diff(sc) <= 0).all()
            assert int(ids[0]) == int(oracle_planted(go, b))
            assert np.array_equal(rb4.row(b)[0], ids) and np.array_equal(rb4.row(b)[1], sc)
            for i, s in zip(ids, sc):                       # reported scores are the oracle's, bit for bit
                assert oracle.cosine(Q[b], oracle.gen_rows(go, int(i), 1, d)[0]) == s
        # ONE exhaustive oracle scan of all 10M rows for every query: ids and scores bit-equal
        for b, (ei, es) in enumerate(exhaustive_topk(oracle, idx, go, Q, 10)):
            assert np.array_equal(r[b].row(0)[0], ei), (b, r[b].row(0)[0], ei)
            assert np.array_equal(r[b].row(0)[1].view(np.uint64), es.view(np.uint64)), b
        # deep_search shape on top: RRF of the exact vector stage with a keyword list
        kw = [int(x) for x in r[0].row(0)[0][:3]] + [123, 9_999_999, 5_000_000, 77, 4_242_424, 31337, 2]
        f = idx.hybrid(Q[0:1], rb.hybrid_opts(10, 10, 0.3, path=native.PATH_STREAM), [kw]).row(0)
        ek, es, esrc, _ = oracle.rrf(f["vec_ids"], kw)
        assert np.array_equal(f["keys"], ek) and np.array_equal(f["scores"], es) and np.array_equal(f["source"], esrc)


def test_c4_memory_rag_three_lists_full_size(rb, native, oracle):
    """C4: 2M memory rows + 5M doc rows in ONE matrix, batch 256, vector + keyword + freshness lists fused
    (north-star extension, SURVEY N-c4) on the tensor path; the reference-faithful two-list result is the
    fresh_limit=0 special case."""
    n, d, B, now = 7_000_000, 1536, 256, 1_760_000_000_000
    go, gn = gen(oracle, native, n, memory_rows=2_000_000, now_ms=now)
    with rb.VectorIndex(d, n, shadow="f16") as idx:
        idx.generate(gn, n)
        Q = idx.generate_queries(gn, 0, B)
        top = idx.query(Q, 10)                                                   # AUTO → tensor path
        assert top.certified.all()
        # exhaustive oracle scan of all 7M rows for a sample of the batch: the tensor path's ids and scores are
        # the oracle's, bit for bit (first pass certified by the rigorous bound, or escalated)
        sub = np.arange(0, B, 32)
        for b, (ei, es) in zip(sub, exhaustive_topk(oracle, idx, go, Q[sub], 10)):
            assert np.array_equal(top.row(b)[0], ei), (b, top.row(b)[0], ei)
            assert np.array_equal(top.row(b)[1].view(np.uint64), es.view(np.uint64)), b
        first = idx.query(Q, 10, flags=native.SEARCH_NO_ESCALATE)
        print(f"C4 first-pass certification (rigorous bound): {int(first.certified.sum())}/{B}")
        kw = [[int(x) for x in top.ids[b, :3]] + [(b * 7919 + j * 104729) % n for j in range(7)] for b in range(B)]
        two = idx.hybrid(Q, rb.hybrid_opts(10, 10, 0.3), kw)
        three = idx.hybrid(Q, rb.hybrid_opts(10, 10, 0.3, fresh_limit=10, fresh_weight=1.0, now_ms=now), kw)
        n_fresh_both = 0
        for b in range(0, B, 8):
            g2, g3 = two.row(b), three.row(b)
            fi, _ = oracle.filter_min_score(*top.row(b), 0.3)                    # top.row(b) == the oracle's (above)
            assert np.array_equal(g2["vec_ids"], fi) and np.array_equal(g3["vec_ids"], fi)
            ek, es, esrc, _ = oracle.rrf(g2["vec_ids"], kw[b])                    # reference-faithful two-list fusion
            assert np.array_equal(g2["keys"], ek) and np.array_equal(g2["scores"], es) and np.array_equal(g2["source"], esrc)
            vi = [int(i) for i in g3["vec_ids"]]
            mem = [i for i in vi if i < 2_000_000]
            ct, cf, ac, la = zip(*[[x[0] for x in oracle.gen_meta(go, i, 1)] for i in mem]) if mem else ((), (), (), ())
            fr = {i: oracle.freshness(float(c), int(a), int(l), now) for i, c, a, l in zip(mem, cf, ac, la)}
            fl = sorted(mem, key=lambda i: (-fr[i], i))[:10]
            ek3, es3, esrc3, _ = oracle.rrf(vi, kw[b], vec_ctype=[1 if i < 2_000_000 else 0 for i in vi], fresh_keys=fl, fresh_weight=1.0)
            assert np.array_equal(g3["keys"], ek3) and np.array_equal(g3["scores"], es3) and np.array_equal(g3["source"], esrc3)
            n_fresh_both += len(mem)
        assert n_fresh_both > 0                                                  # the freshness list was exercised


def test_c5_shard_size_tensor_path_properties(rb, native, oracle):
    """C5: one rank's shard of the 8-GPU configuration (50M/8 = 6.25M x 1536 bf16, batch 1024, tcgen05 path).
    Every query certified, the planted row wins, reported scores are the oracle's bit for bit, the stream path
    (different kernel, fp32 selection) returns the same ids and scores, idempotent — and an exhaustive oracle scan
    of the whole shard for a sample of the batch."""
    n, d, B = 6_250_000, 1536, 1024
    go, gn = gen(oracle, native, n)
    with rb.VectorIndex(d, n, dtype=native.BF16) as idx:
        idx.generate(gn, n)
        Q = idx.generate_queries(gn, 0, B)
        r = idx.query(Q, 10, path=native.PATH_TENSOR)
        assert r.certified.all() and (r.counts == 10).all()
        assert (np.diff(r.scores, axis=1) <= 0).all()
        for b in range(0, B, 64):
            ids, sc = r.row(b)
            assert int(ids[0]) == int(oracle_planted(go, b))
            for i, s in zip(ids[:3], sc[:3]):
                assert oracle.cosine(Q[b], oracle.gen_rows(go, int(i), 1, d, dtype=oracle.BF16)[0]) == s
        sub = np.arange(0, B, 128)
        rs = idx.query(Q[sub], 10, path=native.PATH_STREAM)
        assert np.array_equal(rs.ids, r.ids[sub]) and np.array_equal(rs.scores, r.scores[sub])
        # exhaustive oracle scan of all 6.25M bf16 rows for those queries: ids and scores bit-equal
        for b, (ei, es) in zip(sub, exhaustive_topk(oracle, idx, go, Q[sub], 10)):
            assert np.array_equal(r.row(b)[0], ei), (b, r.row(b)[0], ei)
            assert np.array_equal(r.row(b)[1].view(np.uint64), es.view(np.uint64)), b
        first = idx.query(Q, 10, path=native.PATH_TENSOR, flags=native.SEARCH_NO_ESCALATE)
        print(f"C5 shard first-pass certification (rigorous bound): {int(first.certified.sum())}/{B}")
        r2 = idx.query(Q, 10, path=native.PATH_TENSOR)
        assert np.array_equal(r.ids, r2.ids) and np.array_equal(r.scores, r2.scores)


def test_wide_rows_and_long_lists_fall_back_gracefully(rb, native, oracle):
    """dim 4096 with K'=70: the 8-query kernel does not fit in shared memory; the library picks a smaller
    grouping instead of failing, and results stay exact."""
    rng = np.random.default_rng(12)
    n, d = 600, 4096
    X = rng.standard_normal((n, d)).astype(np.float32)
    Q = (X[[3, 77, 500, 12, 13, 14]] + 0.3 * rng.standard_normal((6, d))).astype(np.float32)
    with rb.VectorIndex(d, n) as idx:
        idx.upload(X)
        r = idx.query(Q, 64, path=native.PATH_STREAM)
        for b in range(6):
            ei, es = oracle.topk(X, Q[b], 64)
            assert np.array_equal(r.row(b)[0], ei) and np.array_equal(r.row(b)[1], es)
    with pytest.raises(rb.RagError):
        rb.VectorIndex(8193, 10)


def test_c_abi_from_plain_c(rb, native, oracle, tmp_path):
    """tests/c/abi_demo.c: the boundary used from C (what the N-API shim does), checked against the oracle."""
    import subprocess

    exe = str(tmp_path / "abi_demo")
    lib_dir = os.path.join(ROOT, "rag_era_b200")
    subprocess.run(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "abi_demo.c"), "-o", exe,
                    "-L", lib_dir, "-lragera", f"-Wl,-rpath,{lib_dir}", "-lpthread"], check=True)
    n = 20000
    out = subprocess.run([exe, str(n)], capture_output=True, text=True, check=True).stdout.split("\n")
    lines = out[:3]
    assert out[3:6] == lines                      # the same requests through rag_batcher_submit_async + callback
    go = oracle.make_gen(n, n_clusters=32)
    X = oracle.gen_rows(go, 0, n, 256)
    Q = oracle.gen_queries(go, 0, 3, 256)
    kws = [[1, 2, 3, 4], [], [n - 1, 7]]
    for b, line in enumerate(lines):
        _, used, cnt, key0, score0, cert = line.split()
        e = oracle.hybrid_search(X, Q[b], 10, 0.3, kws[b])
        assert int(used) == int(e["used_rrf"]) and int(cnt) == len(e["keys"]) and int(cert) == 1
        assert int(key0) == int(e["keys"][0]) and float.fromhex(score0) == float(e["scores"][0])

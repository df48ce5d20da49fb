"""The oracle against the known-answer vectors (CPU only).

PARITY UNPINNED: the reference ships no tests or golden vectors for this path.
The vectors here are (a) tests/golden/kat_rrf.json — a statement-by-statement Python
transliteration of the cited reference lines, and (b) SURVEY.md Appendix A's hand-derived
hex floats. Two independent restatements + the C oracle must agree bit for bit.
"""
import math
import struct

import numpy as np
import pytest

CT = {"document": 0, "memory": 1, "code": 2}
SRC = {0: "vector", 1: "keyword", 2: "both", 3: "freshness"}


def ulp_diff(a: float, b: float) -> int:
    ia, ib = struct.unpack("<q", struct.pack("<d", a))[0], struct.unpack("<q", struct.pack("<d", b))[0]
    return abs(ia - ib)


def test_rrf_golden_bit_exact(oracle, golden):
    assert len(golden["rrf"]) >= 48
    for case in golden["rrf"]:
        c = case["config"]
        cfg = oracle.RRFConfig(c["k"], c["vectorWeight"], c["keywordWeight"], c["bothBonus"])
        vk = [v[0] for v in case["vector"]]
        vt = [CT[v[1]] for v in case["vector"]]
        keys, scores, src, ct = oracle.rrf(vk, case["keyword"], cfg, vec_ctype=vt)
        exp = case["expect"]
        assert len(keys) == len(exp), case["name"]
        for i, e in enumerate(exp):
            assert int(keys[i]) == e["key"], (case["name"], i)
            assert float(scores[i]).hex() == float.fromhex(e["hex"]).hex(), (case["name"], i)
            assert SRC[int(src[i])] == e["source"], (case["name"], i)
            assert int(ct[i]) == CT[e["contentType"]], (case["name"], i)


def test_rrf_appendix_a_hex(oracle):
    """SURVEY Appendix A KAT 1-5 (src/lib/hybrid-search.ts:134-202), keys A..D = 1..4."""
    doc = oracle.RRFConfig(60, 1, 1, 0.1)
    code = oracle.RRFConfig(40, 1, 1.3, 0.15)
    k, s, src, _ = oracle.rrf([1, 2, 3], [2, 4], doc)
    assert list(k) == [2, 1, 4, 3] and [SRC[int(x)] for x in src] == ["both", "vector", "keyword", "vector"]
    assert [float(x).hex() for x in s] == [float.fromhex(h).hex() for h in
                                           ("0x1.17a313935f633p-5", "0x1.0c9714fbcda3bp-6", "0x1.0842108421084p-6", "0x1.0410410410410p-6")]
    k, s, src, _ = oracle.rrf([1], [2], doc)                       # KAT2: tie → insertion order
    assert list(k) == [1, 2] and s[0] == s[1] == 1 / 61
    k, s, src, _ = oracle.rrf([1, 2], [2, 1], code)                # KAT3: e + (s + bonus*e)
    assert list(k) == [2, 1]
    assert float(s[0]).hex() == float.fromhex("0x1.e40d151c3f08ap-5").hex()
    assert float(s[1]).hex() == float.fromhex("0x1.e3566759f8bedp-5").hex()
    k, s, src, _ = oracle.rrf([10, 11, 12, 13, 14], [12, 21, 22, 23, 10], doc)   # KAT4
    assert list(k) == [12, 10, 11, 21, 22, 13, 23, 14]
    assert float(s[0]).hex() == float.fromhex("0x1.15547b0cefc26p-5").hex()
    assert float(s[1]).hex() == float.fromhex("0x1.11c15f3bb8fa8p-5").hex()
    assert float(s[7]).hex() == float.fromhex("0x1.f81f81f81f820p-7").hex()
    k, s, src, _ = oracle.rrf([1, 1], [], doc)                     # KAT5: duplicate key in the vector pass
    assert list(k) == [1] and SRC[int(src[0])] == "both"
    assert float(s[0]).hex() == float.fromhex("0x1.0a6c92bff7560p-5").hex()


def test_rrf3_reduces_to_rrf(oracle):
    rng = np.random.default_rng(7)
    for _ in range(50):
        vk = rng.integers(1, 30, rng.integers(0, 29)).tolist()
        kk = rng.integers(1, 30, rng.integers(0, 19)).tolist()
        a = oracle.rrf(vk, kk)
        b = oracle.rrf(vk, kk, fresh_keys=[], fresh_weight=1.0)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)


def test_filter_and_blend_and_freshness(oracle, golden):
    f = golden["filter"]
    ids, sc = oracle.filter_min_score(np.arange(len(f["scores"])), f["scores"], f["min"])
    assert list(ids) == f["kept"] == [0, 1, 2]                     # `<` drops, `>=` keeps (:310)
    for e in golden["freshness"]:
        v = oracle.freshness(e["confidence"], e["accessCount"], e["lastAccessedMs"], e["nowMs"])
        assert ulp_diff(v, float.fromhex(e["hex"])) <= 2
    # Appendix A freshness values
    now = 1_760_000_000_000
    assert oracle.freshness(0.8, 0, now, now) == 0.8
    assert abs(oracle.freshness(0.8, 3, now - 14 * 3600000, now) - 0.45234131555001067) < 1e-15
    assert oracle.freshness(0.9, 10, now - 3600000, now) == 1.0
    # blend (store.ts:160) through memory_rank on single memory hits
    for e in golden["blend"]:
        # choose metadata whose freshness is irrelevant: check the blend arithmetic via a direct formula
        assert float(e["cos"] * 0.7 + e["fresh"] * 0.3).hex() == float.fromhex(e["hex"]).hex()
    cos = [0.9, 0.6, 0.49, 0.8]
    ism = [1, 1, 1, 0]
    conf, acc, last = [0.8, 0.9, 1.0, 1.0], [0, 10, 5, 5], [now, now - 3600000, now, now]
    oi, osc, ofr = oracle.memory_rank(cos, ism, conf, acc, last, now, limit=10, min_relevance=0.5)
    assert list(oi) == [0, 1]                                      # 0.49 < 0.5 dropped, non-memory dropped
    assert osc[0] == 0.9 * 0.7 + 0.8 * 0.3 and osc[1] == 0.6 * 0.7 + 1.0 * 0.3


def test_cosine_topk_semantics(oracle):
    """Upstream-recalled llamaindex semantics (SURVEY KAT 9-12)."""
    rng = np.random.default_rng(3)
    X = rng.standard_normal((200, 96)).astype(np.float32)
    X[17] = X[5]                       # identical rows: earlier row first (KAT 9)
    q = (X[5] + 0.1 * rng.standard_normal(96)).astype(np.float32)
    ids, sc = oracle.topk(X, q, 5)
    assert list(ids[:2]) == [5, 17] and sc[0] == sc[1]
    ids_f, sc_f = oracle.topk(X, q, 5, faithful_sort=True)
    assert np.array_equal(ids, ids_f) and np.array_equal(sc, sc_f)
    ids, sc = oracle.topk(X[:3], q, 10)          # N < k → N results (KAT 10)
    assert len(ids) == 3
    s1 = oracle.cosine(q, X[9])
    s2 = oracle.cosine(q, (X[9] * 4.0).astype(np.float32))          # exact power-of-two scaling (KAT 11)
    assert abs(s1 - s2) < 1e-14
    assert abs(oracle.cosine(X[4], X[4]) - 1.0) < 1e-15            # KAT 12
    # independent numpy restatement of the three sequential sums
    qd, xd = q.astype(np.float64), X[9].astype(np.float64)
    dot = nq = nx = 0.0
    for a, b in zip(qd, xd):
        dot += a * b
    for a in qd:
        nq += a * a
    for b in xd:
        nx += b * b
    assert s1 == dot / (math.sqrt(nq) * math.sqrt(nx))


def test_cosine_topk_golden_bit_exact(oracle, golden):
    """The C oracle against the independent Python transliteration of llamaindex's similarity / getTopKEmbeddings
    (tests/golden/make_kat.py): ids, order under exact ties, and every similarity bit for bit."""
    assert len(golden["cosine_topk"]) >= 8
    for c in golden["cosine_topk"]:
        X = np.array([[float.fromhex(v) for v in r] for r in c["rows"]], dtype=np.float64)
        q = np.array([float.fromhex(v) for v in c["query"]], dtype=np.float64)
        assert np.array_equal(X, X.astype(np.float32)) and np.array_equal(q, q.astype(np.float32))   # fp32-representable
        for faithful in (False, True):
            ids, sc = oracle.topk(X.astype(np.float32), q.astype(np.float32), c["k"], faithful_sort=faithful)
            assert [int(i) for i in ids] == c["ids"], c["name"]
            assert [float(s).hex() for s in sc] == c["similarities"], c["name"]
        for i, h in zip(c["ids"], c["similarities"]):
            assert float(oracle.cosine(q.astype(np.float32), X[i].astype(np.float32))).hex() == h


def test_threads_do_not_change_bits(oracle):
    rng = np.random.default_rng(11)
    X = rng.standard_normal((3000, 64)).astype(np.float32)
    q = rng.standard_normal(64).astype(np.float32)
    a = oracle.topk(X, q, 29, threads=1)
    b = oracle.topk(X, q, 29, threads=8)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_generated_rows_match_in_memory(oracle):
    g = oracle.make_gen(5000, n_clusters=16, dup_period=7, memory_rows=100)
    X = oracle.gen_rows(g, 0, 5000, 64)
    assert np.array_equal(X[6], X[5]) and np.array_equal(X[13], X[12])     # dup_period ties
    q = oracle.gen_queries(g, 0, 3, 64)
    for b in range(3):
        a = oracle.topk(X, q[b], 10)
        c = oracle.topk_generated(g, oracle.F32, 0, 5000, 64, q[b], 10)
        assert np.array_equal(a[0], c[0]) and np.array_equal(a[1], c[1])
    Xb = oracle.gen_rows(g, 0, 100, 64, dtype=oracle.BF16)
    assert np.array_equal(Xb, oracle.f32_to_bf16(X[:100]))
    ct, cf, ac, la = oracle.gen_meta(g, 0, 200)
    assert ct[:100].all() and not ct[100:].any() and (cf >= 0.5).all() and (cf < 1.0).all()


def test_hybrid_branches(oracle):
    """hybridSearch orchestration KATs 6-8 (src/lib/hybrid-search.ts:308-354)."""
    g = oracle.make_gen(2000, n_clusters=8)
    X = oracle.gen_rows(g, 0, 2000, 64)
    q = oracle.gen_queries(g, 0, 1, 64)[0]
    r = oracle.hybrid_search(X, q, 10, 0.3, [])
    assert not r["used_rrf"] and (r["source"] == 0).all() and np.array_equal(r["keys"], r["vec_ids"])
    assert (r["vec_scores"] >= 0.3).all() and np.array_equal(r["scores"], r["vec_scores"])
    kw = [int(r["vec_ids"][1]), 1999, int(r["vec_ids"][0])]
    r2 = oracle.hybrid_search(X, q, 10, 0.3, kw)
    assert r2["used_rrf"] and set(r2["keys"][:2].tolist()) == {int(r["vec_ids"][0]), int(r["vec_ids"][1])}
    assert (np.diff(r2["scores"]) <= 0).all()

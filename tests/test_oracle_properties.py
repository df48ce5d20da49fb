"""Property tests (hypothesis): the C oracle against the independent Python transliteration of the same
reference lines (tests/golden/make_kat.py) on random inputs, plus invariants of the top-k restatement."""
import importlib.util
import os

import numpy as np
from hypothesis import given, settings, strategies as st

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("make_kat", os.path.join(ROOT, "tests", "golden", "make_kat.py"))
make_kat = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_kat)

CT = ["document", "memory", "code"]
SRC = {0: "vector", 1: "keyword", 2: "both"}

keys = st.integers(min_value=1, max_value=25)


@settings(max_examples=300, deadline=None)
@given(vec=st.lists(st.tuples(keys, st.sampled_from(CT)), max_size=29), kw=st.lists(keys, max_size=19),
       k=st.floats(min_value=0.5, max_value=200, allow_nan=False), vw=st.floats(min_value=0.1, max_value=3),
       kww=st.floats(min_value=0.1, max_value=3), bonus=st.floats(min_value=0, max_value=1))
def test_rrf_oracle_equals_python_transliteration(oracle, vec, kw, k, vw, kww, bonus):
    """reciprocalRankFusion (src/lib/hybrid-search.ts:129-208): two independent restatements agree bit for bit."""
    exp = make_kat.reciprocal_rank_fusion(vec, kw, dict(k=k, vectorWeight=vw, keywordWeight=kww, bothBonus=bonus))
    got = oracle.rrf([v[0] for v in vec], kw, oracle.RRFConfig(k, vw, kww, bonus), vec_ctype=[CT.index(v[1]) for v in vec])
    assert [int(x) for x in got[0]] == [e["id"] for e in exp]
    assert [float(x).hex() for x in got[1]] == [float(e["score"]).hex() for e in exp]
    assert [SRC[int(x)] for x in got[2]] == [e["source"] for e in exp]
    assert [CT[int(x)] for x in got[3]] == [e["contentType"] for e in exp]


@settings(max_examples=200, deadline=None)
@given(conf=st.floats(0, 1), acc=st.integers(0, 10_000), hours=st.floats(0, 5000))
def test_freshness_oracle_equals_python(oracle, conf, acc, hours):
    """calculateFreshnessScore (src/lib/memory/freshness.ts:43-55); glibc and CPython share libm here."""
    now = 1_760_000_000_000
    last = now - int(hours * 3600000)
    assert oracle.freshness(conf, acc, last, now) == make_kat.freshness(conf, acc, last, now)


@settings(max_examples=60, deadline=None)
@given(n=st.integers(1, 300), d=st.sampled_from([8, 33, 128]), k=st.integers(1, 40), seed=st.integers(0, 2**31), dup=st.booleans())
def test_topk_is_a_stable_descending_prefix(oracle, n, d, k, seed, dup):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, d)).astype(np.float32)
    if dup and n > 3:
        X[n // 2] = X[0]
        X[n - 1] = X[0]
    q = rng.standard_normal(d).astype(np.float32)
    ids, sc = oracle.topk(X, q, k, faithful_sort=True, threads=1)
    fast_ids, fast_sc = oracle.topk(X, q, k, faithful_sort=False, threads=3)
    assert np.array_equal(ids, fast_ids) and np.array_equal(sc, fast_sc)      # select == full stable sort
    allsc = np.array([oracle.cosine(q, X[i]) for i in range(n)])
    order = sorted(range(n), key=lambda i: (-allsc[i], i))[:k]                 # stable desc, earlier row first
    assert [int(i) for i in ids] == order and np.array_equal(sc, allsc[order])


@settings(max_examples=80, deadline=None)
@given(n=st.integers(1, 250), d=st.sampled_from([8, 64]), k=st.integers(1, 20), seed=st.integers(0, 2**31),
       pick=st.integers(0, 400), dup=st.booleans())
def test_score_filter_commutes_with_topk(oracle, n, d, k, seed, pick, dup):
    """What the library's filter pushdown rests on: hybridSearch takes the best k rows and THEN keeps
    r.score >= minVectorScore (src/lib/hybrid-search.ts:306-317). Dropping the rows below the filter FIRST and taking the
    best k of what is left gives the same list — for a filter placed anywhere, in particular exactly on a row's score and
    one ulp above it — so a row that provably scores below the filter can be ignored by the selection."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, d)).astype(np.float32)
    if dup and n > 2:
        X[n - 1] = X[0]
    q = rng.standard_normal(d).astype(np.float32)
    allsc = np.array([oracle.cosine(q, X[i]) for i in range(n)])
    base = float(allsc[pick % n])
    for m in (base, float(np.nextafter(base, 2.0)), float(np.nextafter(base, -2.0)), -1.5, 1.5):
        ids, sc = oracle.topk(X, q, k)
        a_ids, a_sc = oracle.filter_min_score(ids, sc, m)                       # the reference's order: top-k, then filter
        keep = np.flatnonzero(allsc >= m)                                        # the other order: filter, then top-k
        if len(keep):
            b_ids, b_sc = oracle.topk(np.ascontiguousarray(X[keep]), q, k)
            b_ids = keep[b_ids.astype(np.int64)]
        else:
            b_ids, b_sc = np.zeros(0, np.int64), np.zeros(0)
        assert [int(i) for i in a_ids] == [int(i) for i in b_ids] and np.array_equal(a_sc, b_sc), m

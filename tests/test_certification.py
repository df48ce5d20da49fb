"""Certification of the tensor path: `certified = 1` must be a PROOF that the ids equal an exact scan.

CPU part: the adversarial construction of tests/adversarial.py defeats the round-1 statistical bound and is
covered by the rigorous one (fp64 emulation of the bf16 operands). GPU part: the library itself on that corpus.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import adversarial  # noqa: E402


def tf32_residual(a):
    """Per element: distance from an fp32 value to the FARTHER of its two tf32 neighbours (0 if it is one) — what
    gen.cu's resid2_tf32 uses, valid whether the hardware truncates or rounds to 10 mantissa bits."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    lo = (u & np.uint32(0xFFFFE000)).view(np.float32).astype(np.float64)
    hi = ((u & np.uint32(0xFFFFE000)) + np.uint32(0x2000)).view(np.float32).astype(np.float64)
    ad = a.astype(np.float64)
    return np.where((u & np.uint32(0x1FFF)) == 0, 0.0, np.maximum(np.abs(ad - lo), np.abs(hi - ad)))


def _rig_eps(rho_q, rho_x, ld):
    return rho_q * (1 + rho_x) + rho_x + ld * 2.0 ** -23


def test_adversarial_rows_defeat_the_statistical_bound_but_not_the_rigorous_one():
    X, q, star, _ = adversarial.build()
    exact, approx, rho_q, rho_x = adversarial.emulate(X, q)
    n, k, kp, ld = len(X), 10, 48, 1536
    true_top = np.lexsort((np.arange(n), -exact))[:k]
    cand = np.lexsort((np.arange(n), -approx))[:kp]
    assert true_top[0] == star and star not in cand                  # the best row is not even a candidate
    t, kth = approx[cand[-1]], np.sort(exact[cand])[::-1][k - 1]
    assert kth > t + 0.024 / np.sqrt(ld)                             # ... and the statistical bound certifies anyway
    eps = _rig_eps(rho_q, rho_x, ld)
    assert np.abs(exact - approx).max() <= eps                       # the rigorous bound holds on every row
    assert not kth > t + eps                                         # so it refuses to certify
    assert np.abs(exact - approx).max() > 6 * 0.024 / np.sqrt(ld)    # errors align: 6x the "11 sigma" figure
    # the fp16 shadow of the normalised rows (RAG_INDEX_F16_SHADOW) shrinks the same effect 12x
    e16, a16, rq16, rx16 = adversarial.emulate(X, q, rows="f16")
    assert np.abs(e16 - a16).max() <= rq16 * (1 + rx16) + rx16 and rx16 < 3e-4


def test_rigorous_bound_holds_on_random_and_scaled_rows():
    """|approx - exact| <= rho_q (1 + rho_x) + rho_x on data of very different shapes (Cauchy-Schwarz), for every
    operand kind of the 16-bit tensor path."""
    rng = np.random.default_rng(3)
    for d, scale in [(64, 1.0), (1536, 1e-3), (256, 3e4)]:
        X = (scale * rng.standard_normal((500, d)) * rng.uniform(0.1, 10, (500, 1))).astype(np.float32)
        X[:50, :4] *= 100                                            # outlier dimensions
        for _ in range(20):
            q = (X[rng.integers(0, 500)] + 0.5 * scale * rng.standard_normal(d)).astype(np.float32)
            for rows in ("bf16", "f16", "exact"):
                exact, approx, rho_q, rho_x = adversarial.emulate(X, q, rows)
                assert np.abs(exact - approx).max() <= rho_q * (1 + rho_x) + rho_x + 1e-12


def test_shadow_stream_bound_holds_on_adversarial_and_scaled_rows():
    """RAG_PATH_SHADOW_STREAM: |q.x~/||q|| - cos| <= rho_x (Cauchy-Schwarz; only the rows are rounded) — on the adversarial
    corpus, whose outlier dimensions make fp16 rounding errors align, and on rows of very different scales."""
    X, q, star, _ = adversarial.build()
    exact, approx, rho_x = adversarial.emulate_shadow_stream(X, q)
    assert np.abs(exact - approx).max() <= rho_x + 1e-12 and rho_x < 3e-4
    assert int(np.argmax(approx)) == star                          # 11 significant bits: the best row stays the best
    rng = np.random.default_rng(4)
    for d, scale in [(64, 1.0), (1536, 1e-3), (256, 3e4)]:
        X = (scale * rng.standard_normal((500, d)) * rng.uniform(0.1, 10, (500, 1))).astype(np.float32)
        X[:50, :4] *= 100
        for _ in range(20):
            q = (X[rng.integers(0, 500)] + 0.5 * scale * rng.standard_normal(d)).astype(np.float32)
            exact, approx, rho_x = adversarial.emulate_shadow_stream(X, q)
            assert np.abs(exact - approx).max() <= rho_x + 1e-12
            assert rho_x <= 2.0 ** -11 * 1.01                      # relative fp16 rounding of every element: ||e_x|| <= 2^-11 ||x~||


@pytest.mark.gpu
def test_gpu_shadow_stream_on_the_adversarial_corpus(native, oracle):
    """The library's shadow stream path on the corpus that defeats the statistical tensor bound: the exact answer, the best
    row in first place, and whatever the first pass certifies on its own is the oracle's top-k."""
    import rag_era_b200 as rb

    X, q, star, _ = adversarial.build()
    n, d = X.shape
    rng = np.random.default_rng(12)
    Q = np.stack([q] + [(X[i] + 0.2 * rng.standard_normal(d)).astype(np.float32) for i in rng.integers(0, 2000, 7)])
    with rb.VectorIndex(d, n, shadow="f16") as idx:
        idx.upload(X)
        assert idx.row_residual() < 3e-4
        for b in range(len(Q)):
            ei, es = oracle.topk(X, Q[b], 10)
            r = idx.query(Q[b:b + 1], 10, path=native.PATH_SHADOW_STREAM)
            assert r.certified[0] == 1 and np.array_equal(r.row(0)[0], ei) and np.array_equal(r.row(0)[1], es), b
            raw = idx.query(Q[b:b + 1], 10, path=native.PATH_SHADOW_STREAM, flags=native.SEARCH_NO_ESCALATE)
            if raw.certified[0]:
                assert np.array_equal(raw.row(0)[0], ei) and np.array_equal(raw.row(0)[1], es), b
        assert int(idx.query(Q[:1], 10, path=native.PATH_SHADOW_STREAM).row(0)[0][0]) == star


@pytest.mark.gpu
def test_gpu_adversarial_corpus_rigorous_default_is_exact_and_statistical_is_not(native, oracle):
    import rag_era_b200 as rb

    X, q, star, _ = adversarial.build()
    n, d = X.shape
    rng = np.random.default_rng(11)
    # a batch so that the tensor path is the natural choice: the adversarial query + ordinary ones
    Q = np.stack([q] + [(X[i] + 0.2 * rng.standard_normal(d)).astype(np.float32) for i in rng.integers(0, 2000, 15)])
    with rb.VectorIndex(d, n, bf16_shadow=True) as idx:
        idx.upload(X)
        assert 3.5e-3 < idx.row_residual() < 4.0e-3                  # rho_x: the star row rounds by ~2^-8 relative
        ei, es = oracle.topk(X, q, 10)
        assert int(ei[0]) == star
        # default (rigorous): first pass does not certify the adversarial query, escalation makes it exact
        raw = idx.query(Q, 10, path=native.PATH_TENSOR, flags=native.SEARCH_NO_ESCALATE)
        assert raw.certified[0] == 0 and star not in raw.row(0)[0]
        r = idx.query(Q, 10, path=native.PATH_TENSOR)
        assert r.certified.all()
        for b in range(len(Q)):
            bi, bs = oracle.topk(X, Q[b], 10)
            assert np.array_equal(r.row(b)[0], bi) and np.array_equal(r.row(b)[1], bs), b
        # the round-1 statistical bound certifies the WRONG answer for the adversarial query
        stat = idx.query(Q, 10, path=native.PATH_TENSOR, flags=native.SEARCH_NO_ESCALATE | native.SEARCH_STAT_EPS, slack=38)
        assert stat.certified[0] == 1 and star not in stat.row(0)[0]
        # the measured tensor-path scores (cosine estimates) stay inside the rigorous bound on every row
        S = idx.debug_tensor_scores(Q[:1]).astype(np.float64)[0]
        exact, _, rho_q, rho_x = adversarial.emulate(X, q)
        assert np.abs(S - exact).max() <= _rig_eps(rho_q, rho_x, 1536) + 1e-5
    # the preferred operand, fp16 of the normalised rows: the same corpus certifies in the first pass and is exact
    with rb.VectorIndex(d, n, shadow="f16") as idx:
        idx.upload(X)
        assert idx.row_residual() < 3e-4
        raw = idx.query(Q, 10, path=native.PATH_TENSOR, flags=native.SEARCH_NO_ESCALATE, slack=38)   # K' = 48 as above
        assert raw.certified[0] == 1 and int(raw.row(0)[0][0]) == star
        assert np.array_equal(raw.row(0)[0], ei) and np.array_equal(raw.row(0)[1], es)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["bf16", "f32+bf16", "f32+f16", "tf32"])
def test_gpu_selected_scores_stay_within_the_rigorous_bound(native, oracle, kind):
    """Outlier-dimension rows, every operand kind: max |K2's cosine estimate - exact cosine| <= the bound K4 uses."""
    import rag_era_b200 as rb

    rng = np.random.default_rng(5)
    n, d, B = 3000, 1536, 64
    X = (rng.standard_normal((n, d)) * rng.uniform(0.2, 5, (n, 1))).astype(np.float32)
    X[:, :8] *= rng.choice([1.0, 40.0], (n, 1))                      # half the rows have outlier dimensions
    Q = (X[rng.integers(0, n, B)] + 0.4 * rng.standard_normal((B, d))).astype(np.float32)
    bf = kind == "bf16"
    Xs = oracle.f32_to_bf16(X) if bf else X
    shadow = {"f32+bf16": "bf16", "f32+f16": "f16"}.get(kind)
    with rb.VectorIndex(d, n, dtype=native.BF16 if bf else native.F32, shadow=shadow) as idx:
        idx.upload(Xs)
        S = idx.debug_tensor_scores(Q).astype(np.float64)
        rho_x = idx.row_residual()
        r = idx.query(Q, 10, path=native.PATH_TENSOR)
    Xd = (oracle.bf16_to_f32(Xs) if bf else X).astype(np.float64)
    Qd = Q.astype(np.float64)
    nq = np.linalg.norm(Qd, axis=1)
    exact = (Qd @ Xd.T) / nq[:, None] / np.linalg.norm(Xd, axis=1)[None, :]
    if kind == "tf32":
        err = np.abs(S / nq[:, None] - exact).max(axis=1)            # tf32 keys hold dot/||x||
        rho_q = np.linalg.norm(tf32_residual(Q), axis=1) / nq         # distance to the farther tf32 neighbour
        assert 0 < rho_x <= 2.0 ** -10
    else:
        err = np.abs(S - exact).max(axis=1)                           # 16-bit keys hold the cosine estimate
        Xf = oracle.bf16_to_f32(Xs) if bf else X
        mode = "f16" if kind == "f32+f16" else ("bf16" if kind == "f32+bf16" else "exact")
        rho_q = np.array([adversarial.operands(Xf[:1], Q[b], mode)[2] for b in range(B)])
        want = 0.0 if bf else adversarial.operands(X, Q[0], mode)[3]
        assert abs(rho_x - want) <= 1e-6 + 1e-3 * want, (rho_x, want)  # the device measured what the emulation measures
        assert (rho_q < (4e-4 if mode == "f16" else 3e-3)).all()      # fp16 of q/||q||: 11 significant bits; bf16: 8
    assert (err <= rho_q * (1 + rho_x) + rho_x + d * 2.0 ** -23 + 1e-5).all(), float(err.max())
    assert r.certified.all()
    for b in range(0, B, 7):
        bi, bs = oracle.topk(Xs, Q[b], 10)
        assert np.array_equal(r.row(b)[0], bi) and np.array_equal(r.row(b)[1], bs)


@pytest.mark.gpu
@pytest.mark.parametrize("shadow,dtype_name", [("f16", "f32"), ("bf16", "f32"), (None, "bf16")])
def test_gpu_min_score_filter_pushed_into_selection_and_certification(native, oracle, shadow, dtype_name):
    """hybridSearch keeps r.score >= minVectorScore (hybrid-search.ts:308-314), and filtering commutes with taking the best k.
    The library leans on that twice: the tensor path never makes a row below (min - margin) a candidate, and K4 certifies
    a query whose non-candidates provably lie below the filter whatever its k-th score is. Neither may change a result:
      * min placed EXACTLY on the exact score of the j-th best row (>= keeps it, the next row goes), per query, tensor path;
      * min above every score ("nothing relevant"): empty result, certified in the FIRST pass (no escalation) on the tensor,
        stream and exact paths alike — before, such a query cost an extra corpus pass;
      * one batch, one min: ids / scores / fused order equal the oracle's, every query certified in the first pass."""
    import rag_era_b200 as rb

    n, d, B, k = 30000, 256, 160, 10
    go = oracle.make_gen(n, n_clusters=40, dup_period=0)
    gn = native.GenDesc.from_buffer_copy(bytes(go))
    dt = native.F32 if dtype_name == "f32" else native.BF16
    X = oracle.gen_rows(go, 0, n, d, dtype=oracle.F32 if dt == native.F32 else oracle.BF16)
    with rb.VectorIndex(d, n, dtype=dt, shadow=shadow) as idx:
        idx.generate(gn, n)
        Q = idx.generate_queries(gn, 0, B)
        rng = np.random.default_rng(8)
        kw = [rng.integers(0, n, 4).tolist() for _ in range(B)]
        first_pass = native.SEARCH_NO_ESCALATE
        # the filter exactly on a score
        for b in range(12):
            ei, es = oracle.topk(X, Q[b], k)
            j = int(rng.integers(1, k))
            m = float(es[j])
            e = oracle.hybrid_search(X, Q[b], k, m, kw[b])
            assert len(e["vec_ids"]) >= j + 1                               # rows with the same score as row j stay too
            g = idx.hybrid(Q[b:b + 1], rb.hybrid_opts(k, 4, m, path=native.PATH_TENSOR), [kw[b]]).row(0)
            assert g["certified"] and np.array_equal(g["vec_ids"], e["vec_ids"]) and np.array_equal(g["vec_scores"], e["vec_scores"]), (b, j)
            assert np.array_equal(g["keys"], e["keys"]) and np.array_equal(g["scores"], e["scores"]) and np.array_equal(g["source"], e["source"])
            m_up = float(np.nextafter(es[j], 2.0))                          # one ulp higher: row j goes
            e = oracle.hybrid_search(X, Q[b], k, m_up, kw[b])
            g = idx.hybrid(Q[b:b + 1], rb.hybrid_opts(k, 4, m_up, path=native.PATH_TENSOR), [kw[b]]).row(0)
            assert g["certified"] and np.array_equal(g["vec_ids"], e["vec_ids"]) and len(g["vec_ids"]) < len(ei), (b, j)
        # nothing relevant: certified at once on every path
        for path in (native.PATH_TENSOR, native.PATH_STREAM, native.PATH_EXACT):
            r = idx.hybrid(Q[:32], rb.hybrid_opts(k, 4, 0.999, path=path, flags=first_pass), kw[:32])
            for b in range(32):
                g = r.row(b)
                e = oracle.hybrid_search(X, Q[b], k, 0.999, kw[b])
                assert g["certified"] and len(g["vec_ids"]) == 0, (path, b)
                assert np.array_equal(g["keys"], e["keys"]) and np.array_equal(g["scores"], e["scores"]), (path, b)
        # one batch on the tensor path, min = 0.3 and a min in the middle of the scores: first pass certifies, results exact
        for m in (0.3, float(np.median([oracle.topk(X, Q[b], k)[1][k // 2] for b in range(8)]))):
            r = idx.hybrid(Q, rb.hybrid_opts(k, 4, m, path=native.PATH_TENSOR, flags=first_pass), kw)
            n_cert = 0
            for b in range(B):
                g = r.row(b)
                n_cert += int(g["certified"])
                if not g["certified"]:
                    continue                                                # first pass only: an uncertified query promises nothing
                e = oracle.hybrid_search(X, Q[b], k, m, kw[b])
                assert np.array_equal(g["vec_ids"], e["vec_ids"]) and np.array_equal(g["vec_scores"], e["vec_scores"]), (m, b)
                assert np.array_equal(g["keys"], e["keys"]) and np.array_equal(g["scores"], e["scores"]), (m, b)
            assert n_cert >= B - 2, (m, n_cert)
            full = idx.hybrid(Q, rb.hybrid_opts(k, 4, m, path=native.PATH_TENSOR), kw)       # with escalation: all exact
            for b in range(B):
                e = oracle.hybrid_search(X, Q[b], k, m, kw[b])
                g = full.row(b)
                assert g["certified"] and np.array_equal(g["vec_ids"], e["vec_ids"]) and np.array_equal(g["keys"], e["keys"]) and np.array_equal(g["scores"], e["scores"]), (m, b)
        # queries that resemble nothing in the corpus: (almost) no row clears the floor, the candidate lists stay short or
        # empty — certified in the first pass because whatever was dropped lies below the filter
        Qr = rng.standard_normal((24, d)).astype(np.float32)
        r = idx.hybrid(Qr, rb.hybrid_opts(k, 4, 0.3, path=native.PATH_TENSOR, flags=first_pass), kw[:24])
        for b in range(24):
            g = r.row(b)
            e = oracle.hybrid_search(X, Qr[b], k, 0.3, kw[b])
            assert g["certified"] and np.array_equal(g["vec_ids"], e["vec_ids"]) and np.array_equal(g["vec_scores"], e["vec_scores"]), b
            assert np.array_equal(g["keys"], e["keys"]) and np.array_equal(g["scores"], e["scores"]), b
        # MemoryStore.retrieve's minRelevance is the same kind of filter
        ct, cf, ac, la = oracle.gen_meta(go, 0, n)
        mem = idx.memory_retrieve(Q[:16], 5, 0.999, now_ms=go.now_ms, path=native.PATH_TENSOR)
        assert int(mem["counts"].sum()) == 0

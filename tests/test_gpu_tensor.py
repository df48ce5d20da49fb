"""K2 (tcgen05 tensor path) against numpy / the oracle. Run on a B200: ``pytest -m gpu``."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rb(native):
    import rag_era_b200

    assert native.load().rag_device_count() > 0
    return rag_era_b200


def bf16_round(oracle, a):
    return oracle.bf16_to_f32(oracle.f32_to_bf16(a))


def f16_unit(Q):
    """The tensor path's query operand next to an fp16 shadow: fp16 of q/||q|| (gen.cu::q_to_f16_kernel), as fp64."""
    Qd = Q.astype(np.float64)
    return (Qd / np.linalg.norm(Qd, axis=1, keepdims=True)).astype(np.float32).astype(np.float16).astype(np.float64)


def bf16_unit(oracle, Q):
    """... and next to bf16 rows (a bf16 corpus or a bf16 shadow): bf16 of q/||q|| — kind::f16 takes ONE format."""
    Qd = Q.astype(np.float64)
    return bf16_round(oracle, (Qd / np.linalg.norm(Qd, axis=1, keepdims=True)).astype(np.float32)).astype(np.float64)


@pytest.mark.parametrize("n,d,B", [(256, 256, 1), (300, 64, 5), (1000, 1536, 128), (777, 512, 129), (5000, 256, 256),
                                   (2049, 1024, 300), (70000, 128, 600)])
def test_tensor_scores_match_numpy(rb, native, oracle, n, d, B):
    """The raw K2 scores on a bf16 corpus = <bf16(q/||q||), x> / ||x||, fp32 accumulation."""
    rng = np.random.default_rng(n + d + B)
    X = oracle.f32_to_bf16(rng.standard_normal((n, d)).astype(np.float32))
    Q = rng.standard_normal((B, d)).astype(np.float32)
    with rb.VectorIndex(d, n, dtype=native.BF16) as idx:
        idx.upload(X)
        S = idx.debug_tensor_scores(Q)
    Xf = oracle.bf16_to_f32(X).astype(np.float64)
    E = (bf16_unit(oracle, Q) @ Xf.T) / np.sqrt((Xf * Xf).sum(1))[None, :]
    # the device normalises in fp32 (rsqrtf), numpy in fp64: an element of q/||q|| that sits on a bf16 midpoint may round
    # the other way (2^-8 of ONE element's contribution); a plumbing error would be off by O(0.1)
    assert np.abs(S - E).max() <= 1e-4, float(np.abs(S - E).max())


@pytest.mark.parametrize("n,d,B", [(300, 64, 5), (1000, 1536, 128), (2049, 1024, 300)])
def test_tensor_scores_match_numpy_f16_shadow(rb, native, oracle, n, d, B):
    """fp32 corpus + fp16 shadow of the normalised rows: the accumulator IS the cosine estimate (no epilogue scaling)."""
    rng = np.random.default_rng(n + d + B)
    X = (rng.standard_normal((n, d)) * rng.uniform(0.01, 300, (n, 1))).astype(np.float32)   # row norms over 4 decades
    X[7] = 0                                                                                  # a zero row scores NaN
    Q = (rng.standard_normal((B, d)) * rng.uniform(0.01, 300, (B, 1))).astype(np.float32)
    with rb.VectorIndex(d, n, shadow="f16") as idx:
        idx.upload(X)
        S = idx.debug_tensor_scores(Q)
    Xd = X.astype(np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        Xh = (Xd / np.linalg.norm(Xd, axis=1, keepdims=True)).astype(np.float32).astype(np.float16).astype(np.float64)
    E = f16_unit(Q) @ Xh.T
    ok = np.ones(n, bool)
    ok[7] = False
    assert np.isnan(S[:, 7]).all()
    assert np.abs(S[:, ok] - E[:, ok]).max() <= 5e-5, float(np.abs(S[:, ok] - E[:, ok]).max())


@pytest.mark.parametrize("dtype_name", ["bf16", "f32+shadow", "f32+f16"])
@pytest.mark.parametrize("B,k", [(16, 10), (130, 5), (256, 10), (700, 23)])
def test_tensor_topk_matches_oracle(rb, native, oracle, dtype_name, B, k):
    n, d = 30000, 512
    go = oracle.make_gen(n, n_clusters=64, dup_period=17)
    gn = native.GenDesc.from_buffer_copy(bytes(go))
    bf = dtype_name == "bf16"
    X = oracle.gen_rows(go, 0, n, d, dtype=oracle.BF16 if bf else oracle.F32)
    shadow = None if bf else ("f16" if dtype_name == "f32+f16" else "bf16")
    with rb.VectorIndex(d, n, dtype=native.BF16 if bf else native.F32, shadow=shadow) as idx:
        idx.generate(gn, n)
        Q = idx.generate_queries(gn, 0, B)
        r = idx.query(Q, k, path=native.PATH_TENSOR)
        raw = idx.query(Q, k, path=native.PATH_TENSOR, flags=native.SEARCH_NO_ESCALATE | native.SEARCH_STAT_EPS)
        rig = idx.query(Q, k, path=native.PATH_TENSOR, flags=native.SEARCH_NO_ESCALATE)
        for b in range(0, B, max(1, B // 40)):
            ei, es = oracle.topk(X, Q[b], k)
            gi, gs = r.row(b)
            assert np.array_equal(gi, ei), (b, gi, ei)
            assert np.array_equal(gs, es)
        assert r.certified.all()
        # under the statistical bound the tensor path alone certifies nearly everything on this data; the rigorous
        # bound is wider (measured rounding residuals, no independence assumption; it compensates with a wider
        # candidate window when both operands are rounded); whatever the first pass could not certify was escalated above
        assert raw.certified.mean() > 0.8
        certified = np.flatnonzero(rig.certified)
        for b in certified[:: max(1, len(certified) // 25)]:      # certified without escalation == exact
            ei, es = oracle.topk(X, Q[b], k)
            assert np.array_equal(rig.row(b)[0], ei) and np.array_equal(rig.row(b)[1], es)


def test_tensor_hybrid_batch(rb, native, oracle):
    n, d, B = 20000, 1536, 96
    go = oracle.make_gen(n, n_clusters=32)
    gn = native.GenDesc.from_buffer_copy(bytes(go))
    X = oracle.gen_rows(go, 0, n, d)
    with rb.VectorIndex(d, n, shadow="f16") as idx:
        idx.generate(gn, n)
        Q = idx.generate_queries(gn, 0, B)
        kw = [[int(x) for x in oracle.topk(X, Q[b], 3)[0]] + [n - 1 - b] for b in range(B)]
        o = rb.hybrid_opts(10, 4, 0.3, path=native.PATH_TENSOR)
        res = idx.hybrid(Q, o, kw)
        auto = idx.hybrid(Q, rb.hybrid_opts(10, 4, 0.3), kw)          # AUTO picks the tensor path for batches
        for b in range(B):
            e = oracle.hybrid_search(X, Q[b], 10, 0.3, kw[b])
            g = res.row(b)
            assert np.array_equal(g["keys"], e["keys"]) and np.array_equal(g["scores"], e["scores"])
            assert np.array_equal(g["vec_ids"], e["vec_ids"]) and np.array_equal(g["vec_scores"], e["vec_scores"])
            assert np.array_equal(auto.row(b)["keys"], e["keys"])


def test_bf16_selection_error_is_within_the_stated_tolerance(rb, native, oracle):
    """The stated bf16 tolerance: |cos_bf16 - cos_f64| has sigma ~ 0.0022/sqrt(D) on near-Gaussian
    rows; the STATISTICAL certification bound (RAG_SEARCH_STAT_EPS) is 0.024/sqrt(ld) (~11 sigma). Measure it on
    0.5M (query,row) pairs. (The default bound is the rigorous one: tests/test_certification.py.)"""
    n, d, B = 2000, 1536, 256
    go = oracle.make_gen(n, n_clusters=16)
    gn = native.GenDesc.from_buffer_copy(bytes(go))
    X = oracle.gen_rows(go, 0, n, d).astype(np.float64)
    with rb.VectorIndex(d, n, bf16_shadow=True) as idx:
        idx.generate(gn, n)
        Q = idx.generate_queries(gn, 0, B)
        S = idx.debug_tensor_scores(Q).astype(np.float64)
    Qd = Q.astype(np.float64)
    exact = (Qd @ X.T) / (np.sqrt((Qd * Qd).sum(1))[:, None] * np.sqrt((X * X).sum(1))[None, :])
    err = S - exact                                    # the 16-bit path's keys are cosine estimates
    eps = 0.024 / np.sqrt(d)
    assert np.abs(err).max() < 0.6 * eps, (float(np.abs(err).max()), eps)
    assert err.std() < 0.0022 / np.sqrt(d) * 1.25, float(err.std())
    # the fp16 shadow of the normalised rows: 11 significant bits in both operands, ~6x smaller error
    with rb.VectorIndex(d, n, shadow="f16") as idx:
        idx.generate(gn, n)
        S16 = idx.debug_tensor_scores(Q).astype(np.float64)
    err16 = S16 - exact
    assert np.abs(err16).max() < 0.25 * np.abs(err).max() and err16.std() < 0.25 * err.std(), (float(np.abs(err16).max()), float(err16.std()))


def test_tf32_path_on_fp32_index_without_shadow(rb, native, oracle):
    """An fp32 index without a bf16 shadow is scored by the same tcgen05 kernel in kind::tf32 (TMA rounds the
    fp32 rows and queries to tf32): stated tolerance 0.006/sqrt(ld), ids and scores still the oracle's."""
    n, d, B = 30000, 1536, 200
    go = oracle.make_gen(n, n_clusters=64, dup_period=23)
    gn = native.GenDesc.from_buffer_copy(bytes(go))
    X = oracle.gen_rows(go, 0, n, d)
    with rb.VectorIndex(d, n) as idx:                       # no shadow
        idx.generate(gn, n)
        Q = idx.generate_queries(gn, 0, B)
        S = idx.debug_tensor_scores(Q[:64]).astype(np.float64)
        Xd, Qd = X.astype(np.float64), Q[:64].astype(np.float64)
        exact = (Qd @ Xd.T) / (np.sqrt((Qd * Qd).sum(1))[:, None] * np.sqrt((Xd * Xd).sum(1))[None, :])
        err = S / np.sqrt((Qd * Qd).sum(1))[:, None] - exact
        eps = 0.006 / np.sqrt(d)
        assert np.abs(err).max() < 0.6 * eps, (float(np.abs(err).max()), eps)
        r = idx.query(Q, 10, path=native.PATH_TENSOR)
        auto = idx.query(Q, 10)                              # AUTO: batches of 8+ → tensor (tf32) path
        for b in range(0, B, 5):
            ei, es = oracle.topk(X, Q[b], 10)
            assert np.array_equal(r.row(b)[0], ei) and np.array_equal(r.row(b)[1], es)
            assert np.array_equal(auto.row(b)[0], ei)
        assert r.certified.all()
        small = idx.query(Q[:16], 5, path=native.PATH_TENSOR)   # fewer queries than one 128-row TMA box
        for b in range(16):
            ei, es = oracle.topk(X, Q[b], 5)
            assert np.array_equal(small.row(b)[0], ei) and np.array_equal(small.row(b)[1], es)


def _expected_candidates(scores_row, kp):
    """Top-K' rows of one query by (score desc, row asc); NaN (zero-norm rows) never selected."""
    valid = np.flatnonzero(~np.isnan(scores_row))
    order = valid[np.lexsort((valid, -scores_row[valid].astype(np.float64)))]
    return order[:kp]


@pytest.mark.parametrize("n,d,B,kp", [(70000, 64, 600, 32), (70000, 64, 600, 48), (70000, 64, 256, 8), (5000, 128, 130, 33),
                                      (200, 64, 40, 32), (20, 64, 3, 32), (40000, 64, 1100, 17), (120000, 64, 2050, 32)])
def test_tensor_candidate_lists_are_the_exact_topk_of_the_scores(rb, native, oracle, n, d, B, kp):
    """The fused selection (register-network first tile, threshold votes, window folds, K3 merge) returns
    exactly the K' best (score desc, row asc) of the scores the same launch computed."""
    rng = np.random.default_rng(n * 31 + B + kp)
    X = rng.standard_normal((n, d)).astype(np.float32)
    X[rng.integers(0, n, size=max(1, n // 50))] = 0.0                       # zero-norm rows are never candidates
    X = oracle.f32_to_bf16(X)
    Q = rng.standard_normal((B, d)).astype(np.float32)
    with rb.VectorIndex(d, n, dtype=native.BF16) as idx:
        idx.upload(X)
        S, rows, cs = idx.debug_tensor_candidates(Q, kp)
    for b in range(0, B, 1 if B <= 1100 else 9):   # the largest case (52 tiles per CTA pair) is sampled
        e = _expected_candidates(S[b], kp)
        assert np.array_equal(rows[b, :len(e)], e), (b, rows[b], e)
        assert (rows[b, len(e):] == -1).all()
        assert np.array_equal(cs[b, :len(e)], S[b, e])


@pytest.mark.parametrize("distinct,kp", [(16, 32), (3, 48), (1, 10), (300, 32)])
def test_tensor_candidate_ties_go_to_the_lower_row(rb, native, oracle, distinct, kp):
    """A corpus of a few distinct rows repeated over and over: every score is tied many times; the K' candidates
    are the lowest row ids of the best score classes (the reference's stable sort keeps insertion order)."""
    n, d, B = 30000, 64, 300
    rng = np.random.default_rng(distinct)
    base = oracle.f32_to_bf16(rng.standard_normal((distinct, d)).astype(np.float32))
    X = base[rng.integers(0, distinct, size=n)]
    Q = rng.standard_normal((B, d)).astype(np.float32)
    with rb.VectorIndex(d, n, dtype=native.BF16) as idx:
        idx.upload(X)
        S, rows, cs = idx.debug_tensor_candidates(Q, kp)
    for b in range(0, B, 7):
        e = _expected_candidates(S[b], kp)
        assert np.array_equal(rows[b], e), (b, rows[b], e)


def test_shared_operand_cluster_variant(native):
    """The experimental K2 variant that runs two CTA pairs per cluster and TMA-multicasts the operand they share
    (RAGERA_K2_CLUSTER=1; measured slower than plain pairs on B200 because a 4-CTA cluster strands SMs — DESIGN.md §4 — so
    it is off by default): kept verified here by re-running the selection tests of this file with the switch on."""
    import os
    import subprocess
    import sys

    if os.environ.get("RAGERA_K2_CLUSTER"):
        pytest.skip("already inside the variant run")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_tensor.py"), "-m", "gpu", "-x", "-q",
                        "-k", "topk_matches_oracle or hybrid_batch or ties_go_to or tf32_path or f16_shadow"],
                       env=dict(os.environ, RAGERA_K2_CLUSTER="1"), capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]

"""RAG_PATH_SHADOW_STREAM: one query scored by streaming the fp16 shadow of the NORMALISED rows (K1 at 2 bytes per element)
instead of the fp32 corpus. Selection only — K4 rescored the K' survivors from the fp32 rows in the reference's fp64 order —
so ids and scores must equal the oracle's bit for bit, and `certified = 1` without escalation must be a proof.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rb(native):
    import rag_era_b200

    assert native.load().rag_device_count() > 0, "no GPU visible: the gpu suite needs a B200"
    return rag_era_b200


def check(rb, native, oracle, X, Q, k, path, id_base=0, slack=0):
    with rb.VectorIndex(X.shape[1], len(X), shadow="f16", id_base=id_base) as idx:
        idx.upload(X)
        r = idx.query(Q, k, path=path, slack=slack)
        raw = idx.query(Q, k, path=path, slack=slack, flags=native.SEARCH_NO_ESCALATE)
        for b in range(len(Q)):
            ei, es = oracle.topk(X, Q[b], k, id_base=id_base)
            gi, gs = r.row(b)
            assert np.array_equal(gi, ei), (b, gi, ei)
            assert np.array_equal(gs.view(np.uint64), es.view(np.uint64)), (b, gs, es)
            assert r.certified[b] == 1
            if raw.certified[b]:                     # certified by the first pass alone == exact
                assert np.array_equal(raw.row(b)[0], ei) and np.array_equal(raw.row(b)[1], es), b
        return r, raw


@pytest.mark.parametrize("n,d,k", [(1, 64, 5), (3, 64, 10), (33, 96, 64), (1000, 1536, 10), (5000, 1024, 5), (4097, 100, 12),
                                   (20000, 256, 23), (257, 1536, 2), (40000, 768, 10)])
def test_shadow_stream_topk_matches_oracle(rb, native, oracle, n, d, k):
    rng = np.random.default_rng(n * 17 + d)
    X = rng.standard_normal((n, d)).astype(np.float32) * rng.uniform(0.01, 300.0, (n, 1)).astype(np.float32)   # norms all over
    Q = (X[rng.integers(0, n, 3)] + 0.4 * rng.standard_normal((3, d))).astype(np.float32)
    check(rb, native, oracle, X, Q, k, native.PATH_SHADOW_STREAM, id_base=11)


def test_one_query_on_a_shadowed_index_takes_the_shadow_by_default(rb, native, oracle):
    """AUTO: B == 1 on an fp32 index whose fp16 shadow is larger than L2 (>= 128 MB) streams the shadow (the STREAM class runs
    half as long); results
    are the oracle's either way, through the graph-replayed latency path as well (three calls: capture, replay, replay)."""
    n, d = 100000, 1536
    go = oracle.make_gen(n, n_clusters=64, dup_period=13)
    gn = native.GenDesc.from_buffer_copy(bytes(go))
    X = oracle.gen_rows(go, 0, n, d)
    with rb.VectorIndex(d, n, shadow="f16") as idx:
        idx.generate(gn, n)
        Q = idx.generate_queries(gn, 0, 6)
        for b in range(6):
            kw = [int(i) for i in oracle.topk(X, Q[b], 3)[0]] + [n - 1 - b, 17 + b]
            e = oracle.hybrid_search(X, Q[b], 10, 0.3, kw)
            for path in (native.PATH_AUTO, native.PATH_SHADOW_STREAM, native.PATH_STREAM):
                g = idx.hybrid(Q[b:b + 1], rb.hybrid_opts(10, 5, 0.3, path=path), [kw]).row(0)
                assert np.array_equal(g["vec_ids"], e["vec_ids"]) and np.array_equal(g["vec_scores"], e["vec_scores"]), (b, path)
                assert np.array_equal(g["keys"], e["keys"]) and np.array_equal(g["scores"], e["scores"]) and g["certified"], (b, path)
        # the kernel class that ran: per-launch time of the STREAM class with the shadow vs with the fp32 rows
        per_launch = {}
        idx.profile_enable(True)
        for path in (native.PATH_AUTO, native.PATH_STREAM):
            for _ in range(3):
                idx.query(Q[:1], 10, path=path)
            idx.profile_read()
            for _ in range(20):
                idx.query(Q[:1], 10, path=path)
            ms, cnt = idx.profile_read()["stream"]
            assert cnt == 20
            per_launch[path] = ms / cnt
        idx.profile_enable(False)
        assert per_launch[native.PATH_AUTO] < 0.8 * per_launch[native.PATH_STREAM], per_launch


def test_shadow_stream_exact_ties_and_escalation(rb, native, oracle):
    """Duplicate rows: identical shadow rows, identical scores — ties go to the lower id; more ties than the window holds
    cannot be certified and escalate (shadow -> fp32 stream -> exact) to the oracle's answer."""
    rng = np.random.default_rng(9)
    X = rng.standard_normal((800, 64)).astype(np.float32)
    X[100:180] = X[100]                       # 80 identical rows > K' = 32
    X[300] = X[5] * 3.0                       # same direction, different norm: equal cosine up to rounding
    Q = np.stack([(X[100] + 0.05 * rng.standard_normal(64)).astype(np.float32), (X[5] + 0.01 * rng.standard_normal(64)).astype(np.float32)])
    r, raw = check(rb, native, oracle, X, Q, 10, native.PATH_SHADOW_STREAM)
    assert raw.certified[0] == 0              # flagged, not silently wrong
    ids, sc = r.row(0)
    assert all(int(ids[i]) < int(ids[i + 1]) for i in range(9) if sc[i] == sc[i + 1])


def test_shadow_stream_zero_rows_are_never_selected(rb, native, oracle):
    rng = np.random.default_rng(4)
    X = rng.standard_normal((300, 128)).astype(np.float32)
    X[[0, 7, 299]] = 0.0
    Q = rng.standard_normal((2, 128)).astype(np.float32)
    with rb.VectorIndex(128, 300, shadow="f16") as idx:
        idx.upload(X)
        r = idx.query(Q, 64, path=native.PATH_SHADOW_STREAM)
        for b in range(2):
            assert not set(int(i) for i in r.row(b)[0]) & {0, 7, 299}
        live = np.ones(300, bool)
        live[[0, 7, 299]] = False
        ei, es = oracle.topk(X[live], Q[0], 64)
        assert np.array_equal(np.flatnonzero(live)[ei], r.row(0)[0]) and np.array_equal(es, r.row(0)[1])


def test_shadow_stream_outlier_dimensions_stay_exact(rb, native, oracle):
    """Rows whose energy sits in a few dimensions round worst in fp16; whatever the first pass certifies is exact and the
    measured residual covers every row's error."""
    rng = np.random.default_rng(5)
    n, d = 5000, 1536
    X = (rng.standard_normal((n, d)) * rng.uniform(0.2, 5, (n, 1))).astype(np.float32)
    X[:, :8] *= rng.choice([1.0, 40.0], (n, 1))
    X[::7] = (X[::7] * 1e-4).astype(np.float32)              # tiny rows: the shadow holds the NORMALISED row, no underflow
    Q = (X[rng.integers(0, n, 8)] + 0.4 * rng.standard_normal((8, d))).astype(np.float32)
    with rb.VectorIndex(d, n, shadow="f16") as idx:
        idx.upload(X)
        assert idx.row_residual() < 4e-4
    check(rb, native, oracle, X, Q, 10, native.PATH_SHADOW_STREAM)


def test_shadow_stream_needs_the_f16_shadow(rb, native):
    rng = np.random.default_rng(1)
    X = rng.standard_normal((100, 64)).astype(np.float32)
    for kw in ({}, {"shadow": "bf16"}):
        with rb.VectorIndex(64, 100, **kw) as idx:
            idx.upload(X)
            with pytest.raises(rb.RagError) as ei:
                idx.query(X[:1], 5, path=native.PATH_SHADOW_STREAM)
            assert ei.value.code == native.ERR_UNSUPPORTED
            assert idx.query(X[:1], 5).certified[0] == 1     # AUTO falls back to the fp32 stream

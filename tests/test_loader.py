"""§8f N1 — the reference's persisted index (llamaindex vector_store.json). The parser is host-only,
so its parity with Python's json module is checked on CPU; loading into a device index is a gpu test."""
import json
import os

import numpy as np
import pytest


def write_store(path, ids, X, extra_first=True):
    d = {}
    if extra_first:
        d["textIdToRefDocId"] = {i: "doc-" + i for i in ids}
    d["embeddingDict"] = {i: [float(v) for v in row] for i, row in zip(ids, X)}
    d["metadataDict"] = {i: {"nested": [1, {"a": "}]\\\""}], "type": "memory" if k % 3 == 0 else None} for k, i in enumerate(ids)}
    with open(path, "w") as f:
        json.dump(d, f)


def test_parse_vector_store_json_matches_python_json(native, tmp_path):
    rng = np.random.default_rng(1)
    n, dim = 1000, 48
    X = rng.standard_normal((n, dim)).astype(np.float32)
    X[3, 5] = 1e-30
    X[4, 6] = -123456.789
    ids = [f"a1b2c3d4-{k:04d}-4e5f-8a9b-{k:012x}" for k in range(n)]
    p = str(tmp_path / "vector_store.json")
    write_store(p, ids, X)
    got_ids, got = native.parse_vector_store_json(p, dim, slab_rows=128)
    assert got_ids == ids                                 # file order == insertion order
    assert np.array_equal(got, X)                         # repr(float32 → double) round-trips exactly
    # pretty-printed file, embeddingDict first, values with exponents
    with open(p) as f:
        d = json.load(f)
    with open(p, "w") as f:
        json.dump({"embeddingDict": d["embeddingDict"], "x": [1e5, -2.5E-3, True, None]}, f, indent=2)
    got_ids2, got2 = native.parse_vector_store_json(p, dim)
    assert got_ids2 == ids and np.array_equal(got2, X)


def test_parse_errors(native, tmp_path):
    p = str(tmp_path / "vs.json")
    open(p, "w").write(json.dumps({"embeddingDict": {"a": [1, 2, 3], "b": [1, 2]}}))
    with pytest.raises(native.RagError) as e:
        native.parse_vector_store_json(p, 3)
    assert "has 2 values" in str(e.value)
    open(p, "w").write(json.dumps({"other": 1}))
    with pytest.raises(native.RagError):
        native.parse_vector_store_json(p, 3)
    with pytest.raises(native.RagError):
        native.parse_vector_store_json(str(tmp_path / "missing.json"), 3)
    open(p, "w").write('{"embeddingDict": {"a": [1, 2, 3}')
    with pytest.raises(native.RagError):
        native.parse_vector_store_json(p, 3)


@pytest.mark.gpu
def test_load_vector_store_into_index(native, oracle, tmp_path):
    import rag_era_b200 as rb

    rng = np.random.default_rng(2)
    n, dim = 700, 256
    X = rng.standard_normal((n, dim)).astype(np.float32)
    ids = [f"node-{k}" for k in range(n)]
    p = str(tmp_path / "vector_store.json")
    write_store(p, ids, X)
    with rb.VectorIndex(dim, n + 10) as idx:
        idx.upload(X[:10])                                   # loading appends after existing rows
        got = idx.load_vector_store(p)
        assert got == ids and idx.rows == n + 10
        assert np.array_equal(idx.read_rows(10, n), X)
        q = (X[123] + 0.1 * rng.standard_normal(dim)).astype(np.float32)
        ei, es = oracle.topk(np.vstack([X[:10], X]), q, 5)
        gi, gs = idx.query(q, 5).row(0)
        assert np.array_equal(gi, ei) and np.array_equal(gs, es)
    with rb.VectorIndex(dim, n, dtype=native.BF16) as idx:     # a bf16 index narrows with RNE
        idx.load_vector_store(p)
        assert np.array_equal(idx.read_rows(0, n), oracle.f32_to_bf16(X))

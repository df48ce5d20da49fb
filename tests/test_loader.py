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


# ---- metadataDict (contentType rule of hybrid-search.ts:229-234) and the binary sidecar ---------------------------
def write_store_with_metadata(path, ids, X, shuffle_meta=False):
    """metadataDict as SimpleVectorStore.add writes it: flat node metadata per id, after embeddingDict."""
    meta = {}
    for k, i in enumerate(ids):
        m = {"_node_type": "TextNode", "document_id": "doc-" + i, "nested": {"type": "memory", "x": [1, "]}"]}}
        if k % 5 == 0:
            m.update(type="memory", memoryId=f"mem-{k}", memoryType="preference")
        elif k % 5 == 1:
            m["language"] = "typescript"
        elif k % 5 == 2:
            m["language"] = ""            # `!== undefined` → still code
        elif k % 5 == 3:
            m.update(type="note", memoryId=7)
        meta[i] = m
    if shuffle_meta:
        order = list(meta)
        np.random.default_rng(0).shuffle(order)
        meta = {i: meta[i] for i in order}
        del meta[ids[4]]                   # a node without an entry is a document
    with open(path, "w") as f:
        json.dump({"embeddingDict": {i: [float(v) for v in row] for i, row in zip(ids, X)}, "textIdToRefDocId": {},
                   "metadataDict": meta}, f)


def expected_ctype(native, n):
    ct = np.full(n, native.CT_DOCUMENT, np.uint8)
    ct[0::5] = native.CT_MEMORY
    ct[1::5] = native.CT_CODE
    ct[2::5] = native.CT_CODE
    return ct


@pytest.mark.parametrize("shuffle", [False, True])
def test_parse_vector_store_metadata(native, tmp_path, shuffle):
    rng = np.random.default_rng(3)
    n, dim = 103, 8
    X = rng.standard_normal((n, dim)).astype(np.float32)
    ids = [f"node-{k:03d}" for k in range(n)]
    p = str(tmp_path / "vector_store.json")
    write_store_with_metadata(p, ids, X, shuffle_meta=shuffle)
    ct, mem, found = native.parse_vector_store_metadata(p, ids)
    assert found
    assert np.array_equal(ct, expected_ctype(native, n))
    assert mem == [f"mem-{k}" if k % 5 == 0 else "" for k in range(n)]
    # a store without metadataDict: every row a document
    with open(p, "w") as f:
        json.dump({"embeddingDict": {i: [0.0] * dim for i in ids}}, f)
    ct, mem, found = native.parse_vector_store_metadata(p, ids)
    assert not found and not ct.any() and mem == [""] * n


def make_cache_inputs(n, dim, seed=5):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, dim)).astype(np.float32)
    ids = [f"id-{k}-{'x' * (k % 7)}" for k in range(n)]
    meta = (rng.integers(0, 3, n).astype(np.uint8), rng.random(n), rng.integers(0, 50, n).astype(np.int32),
            rng.integers(1_700_000_000_000, 1_800_000_000_000, n).astype(np.int64))
    keys = rng.integers(0, 2**63, n).astype(np.uint64)
    return X, ids, meta, keys


@pytest.mark.parametrize("n,dim", [(0, 16), (1, 3), (4096, 16), (4097, 24), (10000, 100)])
def test_cache_round_trip_on_the_host(native, tmp_path, n, dim):
    X, ids, meta, keys = make_cache_inputs(n, dim)
    p = str(tmp_path / "kb.ragera")
    native.cache_write_host(p, X, ids=ids, meta=meta, keys=keys)
    info = native.cache_info(p)
    assert (info.rows, info.dim, info.dtype, info.flags) == (n, dim, native.F32, native.CACHE_META | native.CACHE_KEYS)
    assert not os.path.exists(p + ".tmp")
    got = native.cache_read_host(p)
    assert np.array_equal(got["rows"], X) and got["ids"] == ids and np.array_equal(got["keys"], keys)
    for a, b in zip((got["content_type"], got["confidence"], got["access_count"], got["last_access_ms"]), meta):
        assert np.array_equal(a, b)
    if n > 4096:                                              # a shard's row range crosses a checksum block
        part = native.cache_read_host(p, first_row=4000, nrows=n - 4000 - 1)
        assert np.array_equal(part["rows"], X[4000:n - 1]) and np.array_equal(part["keys"], keys[4000:n - 1])
        assert part["ids"] == ids                             # ids come back whole
    # rows only (no metadata, keys, ids), bf16 bits
    bits = (X.view(np.uint32) >> 16).astype(np.uint16)
    native.cache_write_host(p, bits, dtype=native.BF16)
    got = native.cache_read_host(p)
    assert got["info"].dtype == native.BF16 and got["info"].flags == 0 and got["ids"] == []
    assert np.array_equal(got["rows"], bits) and got["content_type"] is None and got["keys"] is None


def test_cache_detects_corruption_and_staleness(native, tmp_path):
    n, dim = 9000, 32
    X, ids, meta, keys = make_cache_inputs(n, dim, seed=6)
    src = str(tmp_path / "vector_store.json")
    open(src, "w").write("{}")
    p = str(tmp_path / "kb.ragera")
    native.cache_write_host(p, X, ids=ids, meta=meta, keys=keys, source_json=src)
    assert native.cache_is_fresh(p, src)
    good = open(p, "rb").read()

    def corrupt(offset, what):
        b = bytearray(good)
        b[offset] ^= 0x40
        open(p, "wb").write(bytes(b))
        with pytest.raises(native.RagError) as e:
            native.cache_read_host(p)
        assert what in str(e.value), str(e.value)

    corrupt(3, "bad magic")
    corrupt(40, "header checksum")
    corrupt(128 + 5000 * dim * 4 + 7, "rows 4096..8192 are corrupt")       # one flipped bit in row 5000
    corrupt(len(good) - 20, "metadata checksum")
    open(p, "wb").write(good[:-16])
    with pytest.raises(native.RagError) as e:
        native.cache_info(p)
    assert "truncated" in str(e.value)
    assert not native.cache_is_fresh(p, src)                                # malformed → not fresh, never an error
    # a row range that avoids the corrupt block still loads (per-block checksums)
    b = bytearray(good)
    b[128 + 5000 * dim * 4 + 7] ^= 0x40
    open(p, "wb").write(bytes(b))
    part = native.cache_read_host(p, first_row=8192, nrows=n - 8192)
    assert np.array_equal(part["rows"], X[8192:])
    # staleness: the JSON grew (index.insert) or was rewritten
    open(p, "wb").write(good)
    assert native.cache_is_fresh(p, src)
    open(src, "a").write(" ")
    assert not native.cache_is_fresh(p, src)
    assert not native.cache_is_fresh(str(tmp_path / "missing.ragera"), src)
    with pytest.raises(native.RagError):
        native.cache_write_host(str(tmp_path / "no_such_dir" / "x.ragera"), X)


@pytest.mark.gpu
def test_open_store_through_the_sidecar(native, oracle, tmp_path):
    import rag_era_b200 as rb

    rng = np.random.default_rng(7)
    n, dim = 5000, 128
    X = rng.standard_normal((n, dim)).astype(np.float32)
    ids = [f"node-{k:04d}" for k in range(n)]
    src = str(tmp_path / "vector_store.json")
    write_store_with_metadata(src, ids, X)
    q = (X[321] + 0.1 * rng.standard_normal(dim)).astype(np.float32)
    ei, es = oracle.topk(X, q, 8)
    with rb.VectorIndex(dim, n) as idx:                                    # cold: parses the JSON, writes the sidecar
        got, hit = idx.open_store(src)
        assert got == ids and not hit and idx.rows == n
        assert np.array_equal(idx.read_row_meta(0, n)["content_type"], expected_ctype(native, n))
        keys = np.arange(n, dtype=np.uint64)[::-1].copy()
        idx.set_row_keys(0, keys)
        idx.set_row_meta(0, expected_ctype(native, n), rng.random(n), np.arange(n, dtype=np.int32), np.full(n, 1_700_000_000_000, np.int64))
        want_meta = idx.read_row_meta(0, n)
        idx.save_cache(src + ".ragera", ids=ids, source_json=src)           # now with Memory columns and fusion keys
    assert native.cache_is_fresh(src + ".ragera", src)
    with rb.VectorIndex(dim, n) as idx:                                    # warm: binary rows straight to HBM
        got, hit = idx.open_store(src)
        assert got == ids and hit and idx.rows == n
        assert np.array_equal(idx.read_rows(0, n), X)
        for k, v in idx.read_row_meta(0, n).items():
            assert np.array_equal(v, want_meta[k]), k
        gi, gs = idx.query(q, 8).row(0)
        assert np.array_equal(gi, ei) and np.array_equal(gs, es)
    with rb.VectorIndex(dim, 3000, id_base=1000) as idx:                   # a shard loads only its own row range
        idx.load_cache(src + ".ragera", first_row=1000, nrows=3000)
        assert idx.rows == 3000 and np.array_equal(idx.read_rows(0, 3000), X[1000:4000])
        assert np.array_equal(idx.read_row_meta(0, 3000)["keys"], want_meta["keys"][1000:4000])
    with rb.VectorIndex(dim, n, dtype=native.BF16) as idx:                 # dtype mismatch: the sidecar is ignored
        got, hit = idx.open_store(src, cache_path=src + ".ragera")
        assert not hit and np.array_equal(idx.read_rows(0, n), oracle.f32_to_bf16(X))
    with rb.VectorIndex(dim, n) as idx:                                    # ... and was rebuilt as bf16; back to fp32 with the
        assert idx.open_store(src)[1] == 0                                 # Memory columns and fusion keys the host had saved
        idx.set_row_keys(0, keys)
        idx.set_row_meta(0, want_meta["content_type"], want_meta["confidence"], want_meta["access_count"], want_meta["last_access_ms"])
        idx.save_cache(src + ".ragera", ids=ids, source_json=src)
    open(src, "a").write("\n")                                             # the JSON changed, but only after the embeddings
    with rb.VectorIndex(dim, n) as idx:
        got, hit = idx.open_store(src)
        assert got == ids and hit == 2                                     # same prefix, nothing appended: re-stamped in place
        assert np.array_equal(idx.read_rows(0, n), X)
    assert native.cache_is_fresh(src + ".ragera", src)
    # index.insert (memory/store.ts:56-67): llamaindex rewrites the JSON with new nodes at the END of embeddingDict
    X2 = rng.standard_normal((7, dim)).astype(np.float32)
    ids2 = [f"memory_{k}" for k in range(7)]
    write_store_with_metadata(src, ids + ids2, np.vstack([X, X2]))
    with rb.VectorIndex(dim, n + 7) as idx:
        got, hit = idx.open_store(src)
        assert got == ids + ids2 and hit == 2 and idx.rows == n + 7        # only the 7 appended embeddings were parsed
        assert np.array_equal(idx.read_rows(0, n + 7), np.vstack([X, X2]))
        meta = idx.read_row_meta(0, n + 7)
        assert np.array_equal(meta["content_type"], expected_ctype(native, n + 7))
        assert np.array_equal(meta["keys"][:n], want_meta["keys"]) and np.array_equal(meta["confidence"][:n], want_meta["confidence"])
        gi, gs = idx.query(q, 8).row(0)
        ei2, es2 = oracle.topk(np.vstack([X, X2]), q, 8)
        assert np.array_equal(gi, ei2) and np.array_equal(gs, es2)
    with rb.VectorIndex(dim, n + 7) as idx:
        assert idx.open_store(src)[1] == 1                                 # and the extended sidecar is fresh
    write_store_with_metadata(src, ids[1:] + ids2, np.vstack([X[1:], X2]))  # a node was removed: the prefix no longer matches
    with rb.VectorIndex(dim, n + 7) as idx:
        got, hit = idx.open_store(src)
        assert got == ids[1:] + ids2 and hit == 0 and np.array_equal(idx.read_rows(0, n + 6), np.vstack([X[1:], X2]))


def test_resume_parses_only_the_appended_embeddings(native, tmp_path):
    rng = np.random.default_rng(3)
    dim = 24
    X = rng.standard_normal((300, dim)).astype(np.float32)
    ids = [f"n{k}" for k in range(300)]
    p = str(tmp_path / "vs.json")
    write_store(p, ids[:200], X[:200], extra_first=False)                    # embeddingDict first, as SimpleVectorStore writes it
    got_ids, got, end = native.parse_vector_store_json(p, dim, resume_offset=0)
    assert got_ids == ids[:200] and np.array_equal(got, X[:200]) and end > 0
    more_ids, more, end2 = native.parse_vector_store_json(p, dim, resume_offset=end)
    assert more_ids == [] and len(more) == 0 and end2 == end                 # nothing after the last embedding
    write_store(p, ids, X, extra_first=False)                                # the same store, 100 nodes appended
    more_ids, more, end3 = native.parse_vector_store_json(p, dim, resume_offset=end, slab_rows=16)
    assert more_ids == ids[200:] and np.array_equal(more, X[200:]) and end3 > end
    assert native.parse_vector_store_json(p, dim, resume_offset=end3)[0] == []
    with pytest.raises(native.RagError):                                     # an offset in the middle of a number
        native.parse_vector_store_json(p, dim, resume_offset=end - 3)
    with pytest.raises(native.RagError):
        native.parse_vector_store_json(p, dim, resume_offset=10**12)


@pytest.mark.parametrize("dtype_name", ["f32", "bf16"])
def test_refresh_routes_fresh_extended_rebuilt(native, oracle, tmp_path, dtype_name):
    """rag_cache_refresh_host: nothing to do / extend in place with the appended embeddings / full rebuild."""
    rng = np.random.default_rng(5)
    dim, n = 32, 9000                                                        # more than two 4096-row checksum blocks
    dt = native.F32 if dtype_name == "f32" else native.BF16
    conv = (lambda a: a) if dt == native.F32 else oracle.f32_to_bf16
    X = rng.standard_normal((n + 500, dim)).astype(np.float32)
    ids = [f"node-{k:05d}" for k in range(n + 500)]
    src = str(tmp_path / "vector_store.json")
    cache = src + ".ragera"
    write_store_with_metadata(src, ids[:n], X[:n])
    assert native.cache_refresh(src, dt, dim) == (0, n)                      # no sidecar yet: full parse
    assert native.cache_refresh(src, dt, dim) == (1, n)                      # fresh
    info = native.cache_info(cache)
    assert info.rows == n and info.source_prefix_bytes > 0
    got = native.cache_read_host(cache)
    assert np.array_equal(got["rows"], conv(X[:n])) and got["ids"] == ids[:n]
    assert np.array_equal(got["content_type"], expected_ctype(native, n))
    for step, m in enumerate((n + 1, n + 130, n + 500)):                     # one memory inserted, then more (block boundaries crossed)
        write_store_with_metadata(src, ids[:m], X[:m])
        assert native.cache_refresh(src, dt, dim) == (2, m), step
        assert native.cache_is_fresh(cache, src)
        got = native.cache_read_host(cache)                                  # every block checksum is verified by the reader
        assert np.array_equal(got["rows"], conv(X[:m])) and got["ids"] == ids[:m]
        assert np.array_equal(got["content_type"], expected_ctype(native, m))
        part = native.cache_read_host(cache, first_row=m - 3, nrows=3)
        assert np.array_equal(part["rows"], conv(X[m - 3:m]))
    assert native.cache_refresh(src, dt, dim) == (1, n + 500)
    # an edit BEFORE the resume point (a node rebuilt): the prefix hash no longer matches → full rebuild
    Y = X[:n + 500].copy()
    Y[17, 3] += 1.0
    write_store_with_metadata(src, ids, Y)
    assert native.cache_refresh(src, dt, dim) == (0, n + 500)
    assert np.array_equal(native.cache_read_host(cache)["rows"], conv(Y))
    # the other dtype: not usable → rebuilt for that dtype
    other = native.BF16 if dt == native.F32 else native.F32
    assert native.cache_refresh(src, other, dim)[0] == 0
    # a malformed appended tail → full parse is attempted and reports the error; no sidecar claims to be fresh
    assert native.cache_refresh(src, dt, dim)[0] == 0
    text = open(src).read()
    cut = text.rindex("]", 0, text.index('"metadataDict"'))
    open(src, "w").write(text[:cut + 1] + ',"broken":[1,2' + text[cut + 1:])
    with pytest.raises(native.RagError):
        native.cache_refresh(src, dt, dim)
    assert not native.cache_is_fresh(cache, src)


def test_append_to_a_100k_row_store_reopens_fast_and_parse_scales(native, tmp_path):
    """one index.insert used to make the whole sidecar stale (full re-parse at 5.8k rows/s on one core).
    Now: (a) the full parse runs on all host threads, (b) after one appended row only that row is parsed and the sidecar is
    extended in place. The store is written by a small C program (tests/c/gen_vector_store.c; Python's json takes minutes)."""
    import shutil
    import subprocess
    import time

    if shutil.which("gcc") is None:
        pytest.skip("needs gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "gen_vs")
    subprocess.run(["gcc", "-O2", "-o", exe, os.path.join(root, "tests", "c", "gen_vector_store.c"), "-lm"], check=True)
    cores = os.cpu_count() or 1
    # (a) throughput at D = 1536
    src = str(tmp_path / "big.json")
    n, dim = 12_000, 1536
    subprocess.run([exe, src, str(n), str(dim), "0"], check=True)
    size_mb = os.path.getsize(src) / 1e6
    t0 = time.perf_counter()
    ids, _ = native.parse_vector_store_json(src, dim, slab_rows=4096, keep_rows=False)
    dt_all = time.perf_counter() - t0
    os.environ["RAGERA_LOADER_THREADS"] = "1"
    try:
        t0 = time.perf_counter()
        native.parse_vector_store_json(src, dim, slab_rows=4096, keep_rows=False)
        dt_one = time.perf_counter() - t0
    finally:
        del os.environ["RAGERA_LOADER_THREADS"]
    rate_all, rate_one = n / dt_all, n / dt_one
    print(f"parse D=1536: {rate_one:.0f} rows/s on 1 thread, {rate_all:.0f} rows/s on {cores} threads "
          f"({size_mb / dt_all:.0f} MB/s of text)")
    assert len(ids) == n
    # throughput thresholds are for a quiet machine (RAGERA_PERF_ASSERTS=1: 71k rows/s measured on 8 cores); on a
    # shared CI box a noisy neighbour must not turn a correctness suite red, so by default the numbers are only printed
    strict = os.environ.get("RAGERA_PERF_ASSERTS") == "1"
    if cores >= 4 and strict:
        assert rate_all > 2.0 * rate_one, (rate_all, rate_one)               # it scales with the cores
        assert rate_all > 3_000 * cores, rate_all                            # ≥ 50k rows/s on the 16-core GPU host
    os.remove(src)
    # (b) 100k rows (D = 256: 100k x 1536 would be a 2 GB file), then ONE memory inserted
    src = str(tmp_path / "store.json")
    n, dim = 100_000, 256
    subprocess.run([exe, src, str(n), str(dim), "0"], check=True)
    t0 = time.perf_counter()
    assert native.cache_refresh(src, native.F32, dim) == (0, n)
    t_full = time.perf_counter() - t0
    subprocess.run([exe, src, str(n), str(dim), "1"], check=True)             # the same store + 1 appended node
    t0 = time.perf_counter()
    assert native.cache_refresh(src, native.F32, dim) == (2, n + 1)
    t_append = time.perf_counter() - t0
    print(f"100k x {dim}: full refresh {t_full:.2f} s, refresh after one appended row {t_append:.3f} s")
    assert t_append < (1.0 if strict else 10.0)        # 0.16 s measured; route 2 above is the structural guarantee
    if strict:
        assert t_append < 0.5 * t_full
    got = native.cache_read_host(src + ".ragera", first_row=n - 1, nrows=2)
    ids_tail, rows_tail, _ = native.parse_vector_store_json(src, dim, resume_offset=native.cache_info(src + ".ragera").source_prefix_bytes)
    assert ids_tail == [] and got["ids"][-1] == f"memory-{n}" and len(got["ids"]) == n + 1
    full_ids, full = native.parse_vector_store_json(src, dim)
    assert np.array_equal(got["rows"], full[n - 1:]) and full_ids == got["ids"]


# ---- the parser against Python's json on arbitrary stores (hypothesis) -------------------------------------------
from hypothesis import HealthCheck, given, settings, strategies as st

_id = st.text(st.characters(blacklist_categories=("Cs",), blacklist_characters="\0"), min_size=1, max_size=12)
_scalar = st.one_of(st.none(), st.booleans(), st.integers(-10**6, 10**6), st.floats(allow_nan=False, allow_infinity=False, width=32),
                    st.text(st.characters(blacklist_categories=("Cs",)), max_size=8))
_json = st.recursive(_scalar, lambda c: st.one_of(st.lists(c, max_size=3), st.dictionaries(st.text(max_size=5), c, max_size=3)), max_leaves=6)


@st.composite
def _stores(draw):
    ids = draw(st.lists(_id, min_size=0, max_size=6, unique=True))
    dim = draw(st.integers(1, 5))
    rows = [[draw(st.floats(allow_nan=False, allow_infinity=False, width=32)) for _ in range(dim)] for _ in ids]
    meta = {}
    for i in draw(st.lists(st.sampled_from(ids), unique=True, max_size=len(ids))) if ids else []:
        m = draw(st.dictionaries(st.sampled_from(["type", "language", "memoryId", "documentName", "x"]), _json, max_size=4))
        meta[i] = m
    extra = draw(st.dictionaries(st.sampled_from(["textIdToRefDocId", "a", "zz"]), _json, max_size=2))
    order = draw(st.permutations(["embeddingDict", "metadataDict"] + sorted(extra)))
    doc = {}
    for k in order:
        doc[k] = {"embeddingDict": dict(zip(ids, rows)), "metadataDict": meta}.get(k, extra.get(k))
    return ids, dim, rows, meta, doc, draw(st.booleans()), draw(st.sampled_from([None, 1, 2]))


@settings(max_examples=150, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(_stores())
def test_parser_agrees_with_python_json_on_arbitrary_stores(native, tmp_path, store):
    ids, dim, rows, meta, doc, ascii_only, indent = store
    p = str(tmp_path / "fuzz.json")
    with open(p, "w", encoding="utf-8") as f:
        json.dump(doc, f, ensure_ascii=ascii_only, indent=indent)      # \\uXXXX escapes (incl. surrogate pairs) or raw UTF-8
    got_ids, X = native.parse_vector_store_json(p, dim, slab_rows=2)
    assert got_ids == ids
    assert np.array_equal(X.view(np.uint32), np.array(rows, np.float32).reshape(len(ids), dim).view(np.uint32))
    ct, mem, found = native.parse_vector_store_metadata(p, ids)
    assert found
    for r, i in enumerate(ids):
        m = meta.get(i, {})
        want = native.CT_MEMORY if m.get("type") == "memory" else native.CT_CODE if "language" in m else native.CT_DOCUMENT
        assert ct[r] == want, (i, m)
        want_mem = m.get("memoryId") if isinstance(m.get("memoryId"), str) else ""
        assert mem[r] == want_mem.replace("\0", "�"), (i, m)


def rows_of(stdout):
    import re

    m = re.search(r"parse rc=0 rows=(\d+)", stdout)
    return int(m.group(1)) if m else 0


def test_loader_and_sidecar_under_sanitizers(tmp_path):
    """loader.cu + store_cache.cu are host-only: built as plain C++ with ASAN + UBSAN (tests/c/host_asan.cc) and driven over good,
    truncated, malformed, unicode and deeply nested stores, then over truncated / bit-flipped sidecars. No sanitizer report, every
    malformed input rejected with a message, no corrupted sidecar ever returns wrong data."""
    import shutil
    import subprocess

    cuda_inc, cuda_lib = "/usr/local/cuda/include", "/usr/local/cuda/lib64"
    if shutil.which("g++") is None or not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA headers")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csrc = os.path.join(root, "rag_era_b200", "csrc")
    exe = str(tmp_path / "host_asan")
    r = subprocess.run(["g++", "-std=c++17", "-g", "-O1", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-x", "c++",
                        "-I" + cuda_inc, "-I" + os.path.join(root, "include"), "-I" + csrc, os.path.join(root, "tests", "c", "host_asan.cc"),
                        os.path.join(csrc, "loader.cu"), os.path.join(csrc, "store_cache.cu"), "-o", exe, "-L" + cuda_lib, "-lcudart",
                        "-Wl,-rpath," + cuda_lib], capture_output=True, text=True)
    if r.returncode != 0 and "sanitize" in r.stderr and "cannot find" in r.stderr:
        pytest.skip("libasan / libubsan are not installed")
    assert r.returncode == 0, r.stderr[-2000:]
    rng = np.random.default_rng(9)
    dim = 6
    ids = [f"id{i}" for i in range(9000)]
    good = json.dumps({"embeddingDict": {i: [float(v) for v in rng.random(dim)] for i in ids},
                       "metadataDict": {i: {"type": "memory", "memoryId": "m" + i, "language": None} for i in ids[::3]}})
    cases = {
        "good": (good, True), "truncated": (good[:len(good) // 2], False), "badnum": (good.replace("0.", "x.", 1), False),
        "badtype": ('{"embeddingDict": {"a": [1,2,3,4,5,6], "b": "oops"}}', False), "empty": ('{"embeddingDict": {}}', True),
        "emptyfile": ("", False),
        "unicode": ('{"embeddingDict": {"\\ud83d\\ude00\\u0000": [1e999,-1e-999,0,1,2,3]}, "metadataDict": {"\\ud83d\\ude00\\u0000": '
                    '{"type": "memory", "memoryId": "\\ud800"}}}', True),
        "dupkeys": ('{"embeddingDict": {"a": [1,1,1,1,1,1]}, "metadataDict": {"zzz": {"type":"memory"}, "a": 5, "a": {"language": {"x": [1,2,{"y":"}"}]}}}}', True),
        "deepnest": ('{"x": ' + "[" * 5000 + "]" * 5000 + ', "embeddingDict": {"a": [1,2,3,4,5,6]}}', True),
    }
    for name, (text, ok) in cases.items():
        p = str(tmp_path / (name + ".json"))
        open(p, "w").write(text)
        r = subprocess.run([exe, p, str(dim), str(tmp_path / (name + ".ragera"))], capture_output=True, text=True, timeout=300,
                           env=dict(os.environ, ASAN_OPTIONS="detect_leaks=1"))
        assert r.returncode == 0, (name, r.stderr[-1500:])
        assert "AddressSanitizer" not in r.stderr and "LeakSanitizer" not in r.stderr and "runtime error" not in r.stderr, (name, r.stderr[-1500:])
        assert ("parse rc=0" in r.stdout) == ok, (name, r.stdout)
        assert "ACCEPTED CORRUPT DATA" not in r.stdout, (name, r.stdout)
        if "forged headers" in r.stdout:
            assert "forged headers: rejected=5 accepted=0" in r.stdout, (name, r.stdout)
        if ok:
            assert "read rc=0 same=1" in r.stdout and "write rc=0 fresh=1" in r.stdout, (name, r.stdout)
            assert "refresh rc=0 route=0" in r.stdout and "refresh rc=0 route=1" in r.stdout and "prefix past the end ok=0" in r.stdout, (name, r.stdout)
            if rows_of(r.stdout) > 0:
                assert "refresh rc=0 route=2 fresh=1" in r.stdout, (name, r.stdout)

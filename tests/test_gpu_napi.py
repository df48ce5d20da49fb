"""§8f N2 — the N-API addon (integration/node/ragera_addon.cc) driven WITHOUT Node: tests/c/napi_mock.cc is a minimal
in-process Node-API host (objects, typed arrays, externals + finalizers, promises, async work on worker threads,
thread-safe functions delivered by its event loop) that plays the
calls of integration/node/native-retrieval.ts. The addon is compiled unmodified against the stub header and linked with
libragera.so; its results must equal the oracle's bit for bit — the same bar as the ctypes path."""
import json
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_host(tmp_path) -> str:
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "napi_mock")
    libdir = os.path.join(ROOT, "rag_era_b200")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-I" + os.path.join(ROOT, "tests", "c", "node_api_stub"),
                        "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "napi_mock.cc"),
                        os.path.join(ROOT, "integration", "node", "ragera_addon.cc"), "-o", exe, "-L" + libdir, "-lragera",
                        "-Wl,-rpath," + libdir, "-lpthread"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def write_input(path, X, Q, k, kw_limit, min_score, ctype, row_keys, kw_keys, kw_counts, conf=None, acc=None, last=None, now_ms=0):
    n = X.shape[0]
    conf = np.zeros(n) if conf is None else conf
    acc = np.zeros(n) if acc is None else acc
    last = np.zeros(n) if last is None else last
    with open(path, "wb") as f:
        f.write(struct.pack("<6Id", n, X.shape[1], Q.shape[0], k, kw_limit, 0, min_score))
        for a, t in ((X, np.float32), (Q, np.float32), (ctype, np.uint8), (row_keys, np.uint64), (kw_keys, np.uint64), (kw_counts, np.uint32),
                     (conf, np.float64), (acc, np.int32), (last, np.int64)):
            f.write(np.ascontiguousarray(a, dtype=t).tobytes())
        f.write(struct.pack("<q", now_ms))


def test_addon_links_and_refuses_without_a_device(native, tmp_path):
    """CPU: the addon + mock host build and link against libragera.so; with no GPU createIndex throws the library's
    RAG_ERR_NO_DEVICE message as a JS exception (there is no fallback to fall back to)."""
    if native.load().rag_device_count() > 0:
        pytest.skip("a GPU is visible: covered by the gpu test")
    exe = build_host(tmp_path)
    inp = str(tmp_path / "in.bin")
    write_input(inp, np.zeros((4, 8), np.float32), np.zeros((1, 8), np.float32), 2, 2, 0.0, np.zeros(4), np.arange(4), np.zeros((1, 2)), np.zeros(1))
    r = subprocess.run([exe, inp], capture_output=True, text=True)
    assert r.returncode == 3 and "createIndex threw: libragera error -7" in r.stderr and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_addon_results_equal_the_oracle(native, oracle, tmp_path):
    exe = build_host(tmp_path)
    rng = np.random.default_rng(17)
    n, d, B, k, kw_limit, min_score = 3000, 128, 6, 10, 5, 0.3
    centres = rng.standard_normal((20, d)).astype(np.float32)
    X = (centres[rng.integers(0, 20, n)] + 0.7 * rng.standard_normal((n, d))).astype(np.float32)
    planted = rng.integers(0, n, B)
    Q = (X[planted] + 0.4 * rng.standard_normal((B, d))).astype(np.float32)
    ctype = (np.arange(n) % 7 == 0).astype(np.uint8)                              # some memory rows
    row_keys = np.arange(n, dtype=np.uint64) // 2 + 1000                            # two rows share a key: duplicates inside the vector list
    kw_keys = np.zeros((B, kw_limit), np.uint64)
    kw_counts = np.array([5, 3, 0, 5, 1, 4], np.uint32)                             # ragged; query 2 has no keyword hits → vector-only branch
    for b in range(B):
        vi, _ = oracle.topk(X, Q[b], k)
        picks = [row_keys[int(vi[0])], 999_999 + b, row_keys[int(vi[min(3, len(vi) - 1)])], 5, row_keys[int(vi[-1])]]
        kw_keys[b, :kw_counts[b]] = picks[:kw_counts[b]]
    now = 1_760_000_000_000
    conf, acc = 0.5 + 0.5 * rng.random(n), rng.integers(0, 40, n).astype(np.int32)
    last = now - rng.integers(0, 72 * 3_600_000, n)
    inp = str(tmp_path / "in.bin")
    write_input(inp, X, Q, k, kw_limit, min_score, ctype, row_keys, kw_keys, kw_counts, conf, acc, last, now)
    r = subprocess.run([exe, inp], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    out = json.loads(r.stdout)
    assert len(out["single"]) == B and len(out["batched"]) == B and len(out["one_call"]) == B and len(out["search"]) == B
    for b in range(B):
        e = oracle.hybrid_search(X, Q[b], k, min_score, kw_keys[b, :kw_counts[b]], row_keys=row_keys, row_ctype=ctype)
        g = out["single"][b]
        assert g["capacity"] == k + kw_limit and g["certified"] == 1
        assert bool(g["usedRrf"]) == e["used_rrf"] == (kw_counts[b] > 0)
        assert g["vecIds"] == [int(i) for i in e["vec_ids"]] and g["vecScores"] == [float(s) for s in e["vec_scores"]]
        assert g["keys"] == [int(x) for x in e["keys"]] and g["scores"] == [float(s) for s in e["scores"]]
        assert g["source"] == [int(s) for s in e["source"]] and g["contentType"] == [int(c) for c in e["ctype"]]
        # the micro-batcher (concurrent submits from worker threads) and the one-call batch give the same answers
        assert out["batched"][b] == g
        assert out["in_flight"][b] == g          # 2B calls in flight at once on one handle: serialised by the library
        assert out["one_call"][b]["keys"] == g["keys"] and out["one_call"][b]["scores"] == g["scores"]
        # the retriever seam (NativeVectorStore.query) and MemoryStore.retrieve's device half (similarityTopK = limit = k)
        t = out["search"][b]
        vi, vs = oracle.topk(X, Q[b], k)
        assert t["ids"] == [int(i) for i in vi] and t["scores"] == [float(x) for x in vs] and t["certified"] == 1
        vj = vi.astype(np.int64)
        oi, osc, ofr = oracle.memory_rank(vs, ctype[vj], conf[vj], acc[vj], last[vj], now, k, min_score)
        assert t["mem_ids"] == [int(vi[i]) for i in oi] and t["mem_relevance"] == [float(vs[i]) for i in oi]
        assert np.allclose(t["mem_scores"], osc, rtol=0, atol=1e-15) and np.allclose(t["mem_freshness"], ofr, rtol=0, atol=1e-15)
    # submit() does not park a pool thread per request: the batched answers above came back through the thread-safe function
    assert out["via_tsfn"] == B
    # 3B submits against a batcher with 16 slots, destroyBatcher called while they are out: every one is answered (overflow
    # through the RAG_ERR_BUSY fallback), a submit after destroyBatcher throws
    assert len(out["overflow"]) == 3 * B and all(out["overflow"][r] == out["single"][r % B] for r in range(3 * B))
    assert "destroyed" in out["closed_batcher"]
    # destroy(handle) with B searches queued on it: they are answered (the index dies with the last of them), a new call throws
    assert out["doomed"] == [out["search"][b]["ids"] for b in range(B)] and "destroyed" in out["closed_index"]
    assert "libragera error" in out["rejected"] and "64" in out["rejected"]        # k beyond RAG_MAX_TOPK rejects the Promise
    assert "Float32Array" in out["thrown"]                                         # a bad argument throws synchronously

"""ctypes binding of the CPU oracle (oracle/oracle.c).

TEST INFRASTRUCTURE ONLY — see oracle/oracle.h. Nothing under ``rag_era_b200``
may import this module; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do.

PARITY UNPINNED: the reference has no tests or golden vectors for this path and
cannot run here (TypeScript, no Node.js); see DESIGN.md §Oracle. Substitute:
``tests/test_reference_pin.py`` pins the restated lines to the mounted reference
source (hashes, constants, and the reference's own arithmetic statements executed
against this oracle).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

F32, BF16 = 0, 1
SRC_VECTOR, SRC_KEYWORD, SRC_BOTH, SRC_FRESHNESS = 0, 1, 2, 3
CT_DOCUMENT, CT_MEMORY, CT_CODE = 0, 1, 2
SOURCE_NAMES = {0: "vector", 1: "keyword", 2: "both", 3: "freshness"}


def build(force: bool = False) -> str:
    """Compile oracle/liboracle.so with the committed Makefile."""
    src = os.path.join(_HERE, "oracle.c")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(p) > os.path.getmtime(_LIB_PATH)
        for p in (src, os.path.join(_HERE, "oracle.h"), os.path.join(_HERE, "..", "include", "ragera_gen.h"))
        if os.path.exists(p)
    )
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _LIB_PATH


class GenDesc(C.Structure):
    """Mirror of rag_gen_desc (include/ragera_gen.h)."""

    _fields_ = [
        ("seed", C.c_uint64),
        ("query_seed", C.c_uint64),
        ("meta_seed", C.c_uint64),
        ("total_rows", C.c_uint64),
        ("n_clusters", C.c_uint32),
        ("noise", C.c_float),
        ("query_noise", C.c_float),
        ("dup_period", C.c_uint32),
        ("memory_rows", C.c_uint64),
        ("now_ms", C.c_int64),
    ]


class RRFConfigC(C.Structure):
    _fields_ = [("k", C.c_double), ("vector_weight", C.c_double),
                ("keyword_weight", C.c_double), ("both_bonus", C.c_double)]


@dataclass
class RRFConfig:
    k: float = 60.0
    vector_weight: float = 1.0
    keyword_weight: float = 1.0
    both_bonus: float = 0.1

    def c(self) -> RRFConfigC:
        return RRFConfigC(self.k, self.vector_weight, self.keyword_weight, self.both_bonus)


def make_gen(total_rows: int, seed: int = 0xC0FFEE, query_seed: int = 0xBEEF, meta_seed: int = 0xF00D,
             n_clusters: int = 4096, noise: float = 0.6, query_noise: float = 0.5, dup_period: int = 0,
             memory_rows: int = 0, now_ms: int = 1_760_000_000_000) -> GenDesc:
    return GenDesc(seed, query_seed, meta_seed, total_rows, n_clusters, noise, query_noise,
                   dup_period, memory_rows, now_ms)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, u64p, f64p, u8p, f32p = C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.POINTER(C.c_uint8), C.POINTER(C.c_float)
        L.oracle_cosine_f32.restype = C.c_double
        L.oracle_cosine_f32.argtypes = [vp, vp, C.c_uint32]
        L.oracle_cosine_bf16.restype = C.c_double
        L.oracle_cosine_bf16.argtypes = [vp, vp, C.c_uint32]
        L.oracle_topk.restype = C.c_int64
        L.oracle_topk.argtypes = [vp, C.c_int, C.c_uint64, C.c_uint32, vp, C.c_uint32, C.c_uint64, vp, vp, C.c_int, C.c_int]
        L.oracle_topk_generated.restype = C.c_int64
        L.oracle_topk_generated.argtypes = [C.POINTER(GenDesc), C.c_int, C.c_uint64, C.c_uint64, C.c_uint32, vp, C.c_uint32, vp, vp, C.c_int, C.c_int]
        L.oracle_filter_min_score.restype = C.c_uint32
        L.oracle_filter_min_score.argtypes = [vp, vp, C.c_uint32, C.c_double]
        L.oracle_rrf.restype = C.c_uint32
        L.oracle_rrf.argtypes = [vp, vp, C.c_uint32, vp, C.c_uint32, C.POINTER(RRFConfigC), vp, vp, vp, vp]
        L.oracle_rrf3.restype = C.c_uint32
        L.oracle_rrf3.argtypes = [vp, vp, C.c_uint32, vp, C.c_uint32, vp, C.c_uint32, C.c_double, C.POINTER(RRFConfigC), vp, vp, vp, vp]
        L.oracle_freshness.restype = C.c_double
        L.oracle_freshness.argtypes = [C.c_double, C.c_int32, C.c_int64, C.c_int64, C.c_double, C.c_double]
        L.oracle_memory_rank.restype = C.c_uint32
        L.oracle_memory_rank.argtypes = [vp, vp, vp, vp, vp, C.c_uint32, C.c_int64, C.c_uint32, C.c_double, vp, vp, vp]
        L.oracle_hybrid_search.restype = C.c_uint32
        L.oracle_hybrid_search.argtypes = [vp, C.c_int, C.c_uint64, C.c_uint32, vp, C.c_uint32, C.c_double, vp, C.c_uint32,
                                           C.POINTER(RRFConfigC), vp, vp, vp, vp, C.POINTER(C.c_uint32), vp, vp, vp, vp,
                                           C.POINTER(C.c_int), C.c_int]
        L.oracle_gen_rows.restype = None
        L.oracle_gen_rows.argtypes = [C.POINTER(GenDesc), C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, vp, C.c_int]
        L.oracle_gen_queries.restype = None
        L.oracle_gen_queries.argtypes = [C.POINTER(GenDesc), C.c_uint64, C.c_uint32, C.c_uint32, vp]
        L.oracle_gen_meta.restype = None
        L.oracle_gen_meta.argtypes = [C.POINTER(GenDesc), C.c_uint64, C.c_uint64, vp, vp, vp, vp]
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _dtype_of(X: np.ndarray) -> int:
    if X.dtype == np.float32:
        return F32
    if X.dtype == np.uint16:
        return BF16
    raise TypeError("corpus must be float32 or uint16 (bf16 bit patterns)")


def default_threads() -> int:
    return max(1, os.cpu_count() or 1)


def cosine(q: np.ndarray, x: np.ndarray) -> float:
    q = np.ascontiguousarray(q, dtype=np.float32)
    x = np.ascontiguousarray(x)
    if _dtype_of(x) == BF16:
        return lib().oracle_cosine_bf16(_p(q), _p(x), q.shape[0])
    return lib().oracle_cosine_f32(_p(q), _p(x), q.shape[0])


def topk(X: np.ndarray, q: np.ndarray, k: int, id_base: int = 0, faithful_sort: bool = False,
         threads: int | None = None):
    """getTopKEmbeddings: returns (ids uint64[m], scores float64[m]), m <= k."""
    X = np.ascontiguousarray(X)
    q = np.ascontiguousarray(q, dtype=np.float32)
    n, d = X.shape
    ids = np.zeros(max(k, 1), dtype=np.uint64)
    sc = np.zeros(max(k, 1), dtype=np.float64)
    m = lib().oracle_topk(_p(X), _dtype_of(X), n, d, _p(q), k, id_base, _p(ids), _p(sc),
                          int(faithful_sort), threads or default_threads())
    return ids[:m].copy(), sc[:m].copy()


def topk_generated(g: GenDesc, dtype: int, row0: int, n: int, d: int, q: np.ndarray, k: int,
                   faithful_sort: bool = False, threads: int | None = None):
    q = np.ascontiguousarray(q, dtype=np.float32)
    ids = np.zeros(max(k, 1), dtype=np.uint64)
    sc = np.zeros(max(k, 1), dtype=np.float64)
    m = lib().oracle_topk_generated(C.byref(g), dtype, row0, n, d, _p(q), k, _p(ids), _p(sc),
                                    int(faithful_sort), threads or default_threads())
    return ids[:m].copy(), sc[:m].copy()


def filter_min_score(ids: np.ndarray, scores: np.ndarray, min_score: float):
    ids = np.array(ids, dtype=np.uint64)
    scores = np.array(scores, dtype=np.float64)
    m = lib().oracle_filter_min_score(_p(ids), _p(scores), len(ids), min_score)
    return ids[:m], scores[:m]


def rrf(vec_keys, kw_keys, cfg: RRFConfig = RRFConfig(), vec_ctype=None, fresh_keys=None,
        fresh_weight: float = 1.0):
    """reciprocalRankFusion on integer keys → (keys, scores, source, ctype)."""
    vk = np.ascontiguousarray(vec_keys, dtype=np.uint64)
    kk = np.ascontiguousarray(kw_keys, dtype=np.uint64)
    fk = np.ascontiguousarray(fresh_keys if fresh_keys is not None else [], dtype=np.uint64)
    vc = np.ascontiguousarray(vec_ctype if vec_ctype is not None else np.zeros(len(vk)), dtype=np.uint8)
    cap = max(1, len(vk) + len(kk) + len(fk))
    ok = np.zeros(cap, dtype=np.uint64)
    os_ = np.zeros(cap, dtype=np.float64)
    src = np.zeros(cap, dtype=np.uint8)
    ct = np.zeros(cap, dtype=np.uint8)
    c = cfg.c()
    if fresh_keys is None:
        n = lib().oracle_rrf(_p(vk), _p(vc), len(vk), _p(kk), len(kk), C.byref(c), _p(ok), _p(os_), _p(src), _p(ct))
    else:
        n = lib().oracle_rrf3(_p(vk), _p(vc), len(vk), _p(kk), len(kk), _p(fk), len(fk), fresh_weight,
                              C.byref(c), _p(ok), _p(os_), _p(src), _p(ct))
    return ok[:n].copy(), os_[:n].copy(), src[:n].copy(), ct[:n].copy()


def freshness(confidence: float, access_count: int, last_access_ms: int, now_ms: int,
              time_decay_factor: float = 0.05, frequency_bonus: float = 0.1) -> float:
    return lib().oracle_freshness(confidence, access_count, last_access_ms, now_ms,
                                  time_decay_factor, frequency_bonus)


def memory_rank(cos, is_memory, confidence, access_count, last_access_ms, now_ms: int,
                limit: int = 10, min_relevance: float = 0.5):
    cos = np.ascontiguousarray(cos, dtype=np.float64)
    n = len(cos)
    im = np.ascontiguousarray(is_memory, dtype=np.uint8)
    cf = np.ascontiguousarray(confidence, dtype=np.float64)
    ac = np.ascontiguousarray(access_count, dtype=np.int32)
    la = np.ascontiguousarray(last_access_ms, dtype=np.int64)
    oi = np.zeros(max(1, n), dtype=np.uint32)
    osc = np.zeros(max(1, n), dtype=np.float64)
    ofr = np.zeros(max(1, n), dtype=np.float64)
    m = lib().oracle_memory_rank(_p(cos), _p(im), _p(cf), _p(ac), _p(la), n, now_ms, limit, min_relevance,
                                 _p(oi), _p(osc), _p(ofr))
    return oi[:m].copy(), osc[:m].copy(), ofr[:m].copy()


def hybrid_search(X: np.ndarray, q: np.ndarray, vector_top_k: int, min_vector_score: float, kw_keys,
                  cfg: RRFConfig = RRFConfig(), row_keys=None, row_ctype=None, threads: int | None = None):
    """hybridSearch (src/lib/hybrid-search.ts:275-355) → dict."""
    X = np.ascontiguousarray(X)
    q = np.ascontiguousarray(q, dtype=np.float32)
    n, d = X.shape
    kk = np.ascontiguousarray(kw_keys, dtype=np.uint64)
    rk = np.ascontiguousarray(row_keys, dtype=np.uint64) if row_keys is not None else None
    rc = np.ascontiguousarray(row_ctype, dtype=np.uint8) if row_ctype is not None else None
    vid = np.zeros(max(1, vector_top_k), dtype=np.uint64)
    vsc = np.zeros(max(1, vector_top_k), dtype=np.float64)
    nvec = C.c_uint32(0)
    cap = max(1, vector_top_k + len(kk))
    ok = np.zeros(cap, dtype=np.uint64)
    os_ = np.zeros(cap, dtype=np.float64)
    src = np.zeros(cap, dtype=np.uint8)
    ct = np.zeros(cap, dtype=np.uint8)
    used = C.c_int(0)
    c = cfg.c()
    m = lib().oracle_hybrid_search(_p(X), _dtype_of(X), n, d, _p(q), vector_top_k, min_vector_score, _p(kk), len(kk),
                                   C.byref(c), _p(rk) if rk is not None else None, _p(rc) if rc is not None else None,
                                   _p(vid), _p(vsc), C.byref(nvec), _p(ok), _p(os_), _p(src), _p(ct), C.byref(used),
                                   threads or default_threads())
    return dict(keys=ok[:m].copy(), scores=os_[:m].copy(), source=src[:m].copy(), ctype=ct[:m].copy(),
                vec_ids=vid[:nvec.value].copy(), vec_scores=vsc[:nvec.value].copy(), used_rrf=bool(used.value))


def gen_rows(g: GenDesc, row0: int, nrows: int, d: int, dtype: int = F32, threads: int | None = None) -> np.ndarray:
    out = np.empty((nrows, d), dtype=np.float32 if dtype == F32 else np.uint16)
    lib().oracle_gen_rows(C.byref(g), row0, nrows, d, dtype, _p(out), threads or default_threads())
    return out


def gen_queries(g: GenDesc, b0: int, nb: int, d: int) -> np.ndarray:
    out = np.empty((nb, d), dtype=np.float32)
    lib().oracle_gen_queries(C.byref(g), b0, nb, d, _p(out))
    return out


def gen_meta(g: GenDesc, row0: int, nrows: int):
    ct = np.empty(nrows, dtype=np.uint8)
    cf = np.empty(nrows, dtype=np.float64)
    ac = np.empty(nrows, dtype=np.int32)
    la = np.empty(nrows, dtype=np.int64)
    lib().oracle_gen_meta(C.byref(g), row0, nrows, _p(ct), _p(cf), _p(ac), _p(la))
    return ct, cf, ac, la


def bf16_to_f32(a: np.ndarray) -> np.ndarray:
    return (a.astype(np.uint32) << 16).view(np.float32)


def f32_to_bf16(a: np.ndarray) -> np.ndarray:
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = u + 0x7FFF + ((u >> 16) & 1)
    return (u >> 16).astype(np.uint16)

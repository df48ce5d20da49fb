"""VectorIndex — the device-resident replacement of llamaindex's ``SimpleVectorStore``.

The reference keeps ``embeddingDict[nodeId] = number[]`` in a JS object and scans it in
insertion order (``src/lib/llm/index-manager.ts:218-227,264-270``; queried through
``index.asRetriever({similarityTopK}).retrieve`` at ``src/lib/hybrid-search.ts:223-224``).
Here the rows live in HBM behind a ``rag_index`` handle; row order = insertion order, and
the chunk id of a row is ``id_base + row`` so that "lower id wins ties" reproduces the
reference's stable sort.

Everything numeric goes through the C ABI (``include/ragera.h``); numpy is used only to
hold the caller's buffers.
"""
from __future__ import annotations

import ctypes as C
import itertools
import threading
from dataclasses import dataclass, field

import numpy as np

from . import _native as N


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@dataclass
class RRFConfig:
    """RRFConfig — src/lib/hybrid-search.ts:40-45."""
    k: float = 60.0
    vector_weight: float = 1.0
    keyword_weight: float = 1.0
    both_bonus: float = 0.1

    def c(self) -> N.RRFConfigC:
        return N.RRFConfigC(self.k, self.vector_weight, self.keyword_weight, self.both_bonus)


@dataclass
class TopK:
    ids: np.ndarray        # uint64 [B, k]
    scores: np.ndarray     # float64 [B, k] exact cosines
    counts: np.ndarray     # uint32 [B]
    certified: np.ndarray  # uint8 [B]

    def row(self, b: int):
        n = int(self.counts[b])
        return self.ids[b, :n], self.scores[b, :n]


@dataclass
class Fused:
    keys: np.ndarray
    scores: np.ndarray
    source: np.ndarray
    content_type: np.ndarray
    counts: np.ndarray
    used_rrf: np.ndarray
    vec_ids: np.ndarray
    vec_scores: np.ndarray
    vec_counts: np.ndarray
    certified: np.ndarray
    _c: N.FusedOut = field(default=None, repr=False)

    def row(self, b: int):
        n = int(self.counts[b])
        return dict(keys=self.keys[b, :n], scores=self.scores[b, :n], source=self.source[b, :n],
                    ctype=self.content_type[b, :n], used_rrf=bool(self.used_rrf[b]),
                    vec_ids=self.vec_ids[b, :int(self.vec_counts[b])],
                    vec_scores=self.vec_scores[b, :int(self.vec_counts[b])], certified=bool(self.certified[b]))


def _alloc_fused(B: int, cap: int, k: int) -> Fused:
    f = Fused(keys=np.empty((B, cap), np.uint64), scores=np.empty((B, cap), np.float64),
              source=np.empty((B, cap), np.uint8), content_type=np.empty((B, cap), np.uint8),
              counts=np.empty(B, np.uint32), used_rrf=np.empty(B, np.uint8),
              vec_ids=np.empty((B, k), np.uint64), vec_scores=np.empty((B, k), np.float64),
              vec_counts=np.empty(B, np.uint32), certified=np.empty(B, np.uint8))
    f._c = N.FusedOut(cap, _ptr(f.keys), _ptr(f.scores), _ptr(f.source), _ptr(f.content_type), _ptr(f.counts),
                      _ptr(f.used_rrf), _ptr(f.vec_ids), _ptr(f.vec_scores), _ptr(f.vec_counts), _ptr(f.certified))
    return f


def hybrid_opts(vector_top_k: int, keyword_limit: int, min_vector_score: float, rrf: RRFConfig = RRFConfig(),
                path: int = N.PATH_AUTO, slack: int = 0, flags: int = 0, fresh_limit: int = 0,
                fresh_weight: float = 1.0, now_ms: int = 0, time_decay_factor: float = 0.0,
                frequency_bonus: float = 0.0, epsilon: float = 0.0) -> N.HybridOpts:
    return N.HybridOpts(vector_top_k, keyword_limit, min_vector_score, rrf.c(), path, slack, flags, fresh_limit,
                        fresh_weight, now_ms, time_decay_factor, frequency_bonus, epsilon)


class VectorIndex:
    """One shard of the chunk-embedding matrix on one GPU."""

    def __init__(self, dim: int, capacity_rows: int, dtype: int = N.F32, device: int = 0,
                 bf16_shadow: bool = False, id_base: int = 0, shadow: str | None = None):
        """``shadow``: 16-bit tensor-path operand kept next to an fp32 corpus — "f16" (fp16 of the normalised rows,
        preferred: tight rigorous certification) or "bf16" (``bf16_shadow=True`` is the same); None: no copy (batches are
        then scored from the fp32 rows as tf32)."""
        self._lib = N.load()
        self.dim, self.capacity_rows, self.dtype, self.device, self.id_base = dim, capacity_rows, dtype, device, id_base
        if shadow not in (None, "f16", "bf16"):
            raise ValueError("shadow must be None, 'f16' or 'bf16'")
        flags = N.INDEX_F16_SHADOW if shadow == "f16" else (N.INDEX_BF16_SHADOW if (shadow == "bf16" or bf16_shadow) else 0)
        desc = N.IndexDesc(capacity_rows, dim, dtype, device, flags, id_base)
        h = C.c_void_p()
        N.check(self._lib.rag_index_create(C.byref(desc), C.byref(h)))
        self._h = h

    # -- lifetime -------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.rag_index_destroy(self._h)
            self._h = None
        for p in getattr(self, "_pinned", []):
            self._lib.rag_host_free(p)
        self._pinned = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def rows(self) -> int:
        return int(self._lib.rag_index_rows(self._h))

    def _np_dtype(self):
        return np.float32 if self.dtype == N.F32 else np.uint16

    # -- rows -----------------------------------------------------------------------------
    def upload(self, rows: np.ndarray, row0: int | None = None):
        """Append (default) or overwrite rows; ``index.insert`` of the reference appends (memory/store.ts:67)."""
        rows = np.ascontiguousarray(rows, dtype=self._np_dtype())
        if rows.ndim != 2 or rows.shape[1] != self.dim:
            raise ValueError(f"rows must be [n, {self.dim}]")
        r0 = self.rows if row0 is None else row0
        N.check(self._lib.rag_index_upload(self._h, r0, rows.shape[0], _ptr(rows)))
        return r0

    def load_vector_store(self, path: str) -> list:
        """Append every embedding of a llamaindex ``vector_store.json`` (the reference's persisted index,
        index-manager.ts:218-220) in file order; returns the node ids, row by row."""
        rows, blob, nbytes = C.c_uint64(0), C.c_void_p(), C.c_uint64(0)
        N.check(self._lib.rag_index_load_vector_store(self._h, path.encode(), C.byref(rows), C.byref(blob), C.byref(nbytes)))
        ids = C.string_at(blob, nbytes.value).decode("utf-8").split("\0")[:-1] if nbytes.value else []
        self._lib.rag_free(blob)
        return ids

    def save_cache(self, cache_path: str, ids=None, source_json: str | None = None):
        """Write the binary sidecar of this handle's rows (+ row metadata / fusion keys if set)."""
        blob = N._ids_blob(ids) if ids is not None else None
        N.check(self._lib.rag_index_save_cache(self._h, cache_path.encode(), blob, len(blob) if blob else 0,
                                               source_json.encode() if source_json else None))

    def load_cache(self, cache_path: str, first_row: int = 0, nrows: int = 0) -> list:
        """Append rows [first_row, first_row+nrows) of a sidecar (nrows=0: to its end); returns ALL node ids of the file."""
        rows, blob, nbytes = C.c_uint64(0), C.c_void_p(), C.c_uint64(0)
        N.check(self._lib.rag_index_load_cache(self._h, cache_path.encode(), first_row, nrows, C.byref(rows), C.byref(blob),
                                               C.byref(nbytes)))
        return N._split_blob(blob, nbytes.value)

    def open_store(self, vector_store_json: str, cache_path: str | None = None):
        """loadIndex (index-manager.ts:246-275) for an empty handle through the binary sidecar. Returns (node ids, route):
        route 1 = the sidecar was fresh, 2 = the JSON had grown by appended embeddings (only those were parsed, the sidecar
        was extended in place), 0 = the JSON was parsed in full (and the sidecar rewritten)."""
        rows, blob, nbytes, hit = C.c_uint64(0), C.c_void_p(), C.c_uint64(0), C.c_int(0)
        N.check(self._lib.rag_index_open_store(self._h, vector_store_json.encode(), cache_path.encode() if cache_path else None,
                                               C.byref(rows), C.byref(blob), C.byref(nbytes), C.byref(hit)))
        return N._split_blob(blob, nbytes.value), int(hit.value)

    def generate(self, gen: N.GenDesc, nrows: int):
        N.check(self._lib.rag_index_generate(self._h, C.byref(gen), nrows))

    def set_row_meta(self, row0: int, content_type=None, confidence=None, access_count=None, last_access_ms=None):
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=t) for a, t in
                ((content_type, np.uint8), (confidence, np.float64), (access_count, np.int32), (last_access_ms, np.int64))]
        n = {len(a) for a in arrs if a is not None}
        if len(n) != 1:
            raise ValueError("metadata arrays must be given and have one common length")
        N.check(self._lib.rag_index_set_row_meta(self._h, row0, n.pop(), *[_ptr(a) for a in arrs]))

    def set_row_keys(self, row0: int, keys):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        N.check(self._lib.rag_index_set_row_keys(self._h, row0, len(keys), _ptr(keys)))

    def read_row_meta(self, row0: int, nrows: int) -> dict:
        """content_type / confidence / access_count / last_access_ms / keys of rows [row0, row0+nrows)."""
        out = dict(content_type=np.empty(nrows, np.uint8), confidence=np.empty(nrows, np.float64),
                   access_count=np.empty(nrows, np.int32), last_access_ms=np.empty(nrows, np.int64),
                   keys=np.empty(nrows, np.uint64))
        N.check(self._lib.rag_index_read_row_meta(self._h, row0, nrows, *[_ptr(a) for a in out.values()]))
        return out

    def read_rows(self, row0: int, nrows: int) -> np.ndarray:
        out = np.empty((nrows, self.dim), dtype=self._np_dtype())
        N.check(self._lib.rag_index_read_rows(self._h, row0, nrows, _ptr(out)))
        return out

    def generate_queries(self, gen: N.GenDesc, b0: int, B: int) -> np.ndarray:
        out = np.empty((B, self.dim), dtype=np.float32)
        N.check(self._lib.rag_generate_queries(self._h, C.byref(gen), b0, B, _ptr(out)))
        return out

    # -- search ---------------------------------------------------------------------------
    def _queries(self, q) -> np.ndarray:
        q = np.ascontiguousarray(q, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"queries must be [B, {self.dim}]")
        return q

    def query(self, queries, similarity_top_k: int, path: int = N.PATH_AUTO, slack: int = 0, flags: int = 0,
              epsilon: float = 0.0) -> TopK:
        """SimpleVectorStore.query → getTopKEmbeddings: rank-ordered (ids, similarities) per query."""
        q = self._queries(queries)
        B, k = q.shape[0], similarity_top_k
        out = TopK(np.empty((B, max(k, 1)), np.uint64), np.empty((B, max(k, 1)), np.float64),
                   np.empty(B, np.uint32), np.empty(B, np.uint8))
        o = N.SearchOpts(k, path, slack, flags, epsilon)
        c = N.TopkOut(_ptr(out.ids), _ptr(out.scores), _ptr(out.counts), _ptr(out.certified))
        N.check(self._lib.rag_search(self._h, _ptr(q), B, C.byref(o), C.byref(c)))
        return out

    @staticmethod
    def _kw(B: int, kw_lists, keyword_limit: int):
        keys = np.zeros((B, max(keyword_limit, 1)), dtype=np.uint64)
        counts = np.zeros(B, dtype=np.uint32)
        if kw_lists is not None:
            for b, lst in enumerate(kw_lists):
                lst = np.asarray(lst, dtype=np.uint64)
                if len(lst) > keyword_limit:
                    raise ValueError("keyword list longer than keyword_limit")
                keys[b, :len(lst)] = lst
                counts[b] = len(lst)
        return keys[:, :keyword_limit].copy() if keyword_limit else keys[:, :0].copy(), counts

    def hybrid(self, queries, opts: N.HybridOpts, kw_lists=None) -> Fused:
        """hybridSearch after the embedding / Meilisearch round trips (hybrid-search.ts:303-354)."""
        q = self._queries(queries)
        B = q.shape[0]
        keys, counts = self._kw(B, kw_lists, opts.keyword_limit)
        out = _alloc_fused(B, max(1, opts.vector_top_k + opts.keyword_limit + opts.fresh_limit), opts.vector_top_k)
        N.check(self._lib.rag_hybrid_search(self._h, _ptr(q), B, C.byref(opts), _ptr(keys), _ptr(counts), C.byref(out._c)))
        return out

    def hybrid_raw(self, q: np.ndarray, opts: N.HybridOpts, kw_keys: np.ndarray, kw_counts: np.ndarray,
                   out: Fused | None = None) -> Fused:
        """rag_hybrid_search on caller-prepared arrays (no per-call list handling): q float32 [B, dim],
        kw_keys uint64 [B, keyword_limit], kw_counts uint32 [B]; ``out`` may be reused across calls."""
        B = q.shape[0]
        if out is None:
            out = _alloc_fused(B, max(1, opts.vector_top_k + opts.keyword_limit + opts.fresh_limit), opts.vector_top_k)
        N.check(self._lib.rag_hybrid_search(self._h, _ptr(q), B, C.byref(opts), _ptr(kw_keys), _ptr(kw_counts), C.byref(out._c)))
        return out

    def alloc_fused(self, B: int, opts: N.HybridOpts) -> Fused:
        return _alloc_fused(B, max(1, opts.vector_top_k + opts.keyword_limit + opts.fresh_limit), opts.vector_top_k)

    def pinned_array(self, shape, dtype) -> np.ndarray:
        """A numpy array over page-locked host memory (rag_host_alloc): H2D/D2H of it are true async DMA.
        The memory lives until the process exits or ``free_pinned`` is called with the array."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = self._lib.rag_host_alloc(max(n, 16))
        if not p:
            raise N.RagError(N.ERR_NOMEM, "rag_host_alloc failed")
        buf = (C.c_uint8 * max(n, 16)).from_address(p)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(p)
        return arr

    def rrf_fuse(self, vec_lists, kw_lists, cfg: RRFConfig = RRFConfig(), vec_ctypes=None) -> Fused:
        """reciprocalRankFusion on integer keys for B independent list pairs (hybrid-search.ts:129-208)."""
        B = len(vec_lists)
        vs = max([len(v) for v in vec_lists] + [1])
        ks = max([len(v) for v in kw_lists] + [1])
        vk, vc = self._kw(B, vec_lists, vs)
        kk, kc = self._kw(B, kw_lists, ks)
        vt = None
        if vec_ctypes is not None:
            vt = np.zeros((B, vs), dtype=np.uint8)
            for b, lst in enumerate(vec_ctypes):
                vt[b, :len(lst)] = lst
        out = _alloc_fused(B, vs + ks, 1)
        c = cfg.c()
        N.check(self._lib.rag_rrf_fuse(self._h, B, C.byref(c), _ptr(vk), _ptr(vt), _ptr(vc), vs, _ptr(kk), _ptr(kc), ks,
                                       C.byref(out._c)))
        return out

    def memory_retrieve(self, queries, limit: int, min_relevance: float = 0.5, now_ms: int = 0,
                        path: int = N.PATH_AUTO, time_decay_factor: float = 0.0, frequency_bonus: float = 0.0,
                        similarity_top_k: int = 0):
        """MemoryStore.retrieve post-processing on the device (memory/store.ts:102-180); similarity_top_k=0 → 2*limit."""
        q = self._queries(queries)
        B = q.shape[0]
        ids = np.empty((B, limit), np.uint64)
        sc, rel, fr = (np.empty((B, limit), np.float64) for _ in range(3))
        cnt = np.empty(B, np.uint32)
        o = N.MemoryOpts(limit, path, min_relevance, now_ms, time_decay_factor, frequency_bonus, similarity_top_k, 0)
        c = N.MemoryOut(_ptr(ids), _ptr(sc), _ptr(rel), _ptr(fr), _ptr(cnt))
        N.check(self._lib.rag_memory_retrieve(self._h, _ptr(q), B, C.byref(o), C.byref(c)))
        return dict(ids=ids, scores=sc, relevance=rel, freshness=fr, counts=cnt)

    def freshness_scores(self, confidence, access_count, last_access_ms, now_ms: int,
                         time_decay_factor: float = 0.0, frequency_bonus: float = 0.0) -> np.ndarray:
        """calculateFreshnessScore over arrays (memory/freshness.ts:37-56)."""
        cf = np.ascontiguousarray(confidence, dtype=np.float64)
        ac = np.ascontiguousarray(access_count, dtype=np.int32)
        la = np.ascontiguousarray(last_access_ms, dtype=np.int64)
        out = np.empty(len(cf), dtype=np.float64)
        N.check(self._lib.rag_freshness_scores(self._h, len(cf), _ptr(cf), _ptr(ac), _ptr(la), now_ms,
                                               time_decay_factor, frequency_bonus, _ptr(out)))
        return out

    # -- staged form + measurement ------------------------------------------------------------
    def stage_batch(self, queries, kw_lists=None, keyword_limit: int = 0):
        q = self._queries(queries)
        keys, counts = self._kw(q.shape[0], kw_lists, keyword_limit)
        N.check(self._lib.rag_stage_batch(self._h, _ptr(q), q.shape[0], _ptr(keys), _ptr(counts), keyword_limit))
        return q.shape[0]

    def stage_window(self, first: int, count: int):
        N.check(self._lib.rag_stage_window(self._h, first, count))

    def hybrid_staged(self, B: int, opts: N.HybridOpts):
        N.check(self._lib.rag_hybrid_search_staged(self._h, B, C.byref(opts)))

    def fetch_fused(self, B: int, opts: N.HybridOpts) -> Fused:
        out = _alloc_fused(B, max(1, opts.vector_top_k + opts.keyword_limit + opts.fresh_limit), opts.vector_top_k)
        N.check(self._lib.rag_fetch_fused(self._h, B, C.byref(opts), C.byref(out._c)))
        return out

    def certified_totals(self) -> tuple[int, int]:
        """(certified, queries) finished by the fusion kernel since the last call (device-side counters)."""
        c, q = C.c_uint64(0), C.c_uint64(0)
        N.check(self._lib.rag_certified_totals(self._h, C.byref(c), C.byref(q)))
        return int(c.value), int(q.value)

    def row_residual(self) -> float:
        """rho_x of the rigorous certification bound (max ||x - operand(x)|| / ||x|| over the loaded rows)."""
        v = float(self._lib.rag_index_row_residual(self._h))
        if v < 0:
            N.check(int(v))
        return v

    def debug_tensor_scores(self, queries) -> np.ndarray:
        q = self._queries(queries)
        out = np.empty((q.shape[0], self.rows), dtype=np.float32)
        N.check(self._lib.rag_debug_tensor_scores(self._h, _ptr(q), q.shape[0], _ptr(out)))
        return out

    def debug_tensor_candidates(self, queries, kp: int):
        """(scores [B][rows] f32, rows [B][kp] int64 (-1 = empty), cand_scores [B][kp] f32) of one K2+K3 pass."""
        q = self._queries(queries)
        scores = np.empty((q.shape[0], self.rows), dtype=np.float32)
        keys = np.zeros((q.shape[0], kp), dtype=np.uint64)
        N.check(self._lib.rag_debug_tensor_candidates(self._h, _ptr(q), q.shape[0], kp, _ptr(scores), _ptr(keys)))
        rows = np.where(keys != 0, (0xFFFFFFFF - (keys & np.uint64(0xFFFFFFFF))).astype(np.int64), -1)
        o = (keys >> np.uint64(32)).astype(np.uint32)
        bits = np.where(o & np.uint32(0x80000000), o ^ np.uint32(0x80000000), ~o).astype(np.uint32)
        return scores, rows, bits.view(np.float32)

    def sync(self):
        N.check(self._lib.rag_sync(self._h))

    def timer_start(self):
        N.check(self._lib.rag_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float(0)
        N.check(self._lib.rag_timer_stop(self._h, C.byref(ms)))
        return float(ms.value)

    @property
    def launch_count(self) -> int:
        return int(self._lib.rag_launch_count(self._h))

    def profile_enable(self, on: bool = True):
        N.check(self._lib.rag_profile_enable(self._h, int(on)))

    def profile_read(self):
        ms = (C.c_float * N.PROF_CLASSES)()
        cnt = (C.c_uint32 * N.PROF_CLASSES)()
        N.check(self._lib.rag_profile_read(self._h, ms, cnt))
        return {N.PROF_NAMES[i]: (float(ms[i]), int(cnt[i])) for i in range(N.PROF_CLASSES)}

    # -- row-sharded multi-GPU ------------------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_uint8 * N.COMM_ID_BYTES)()
        N.check(N.load().rag_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, nranks: int, rank: int, uid: bytes):
        buf = (C.c_uint8 * N.COMM_ID_BYTES).from_buffer_copy(uid)
        N.check(self._lib.rag_comm_init(self._h, nranks, rank, buf))

    def comm_p2p_export(self, nranks: int, rank: int, max_batch: int = 1024, max_k: int = N.MAX_TOPK) -> bytes:
        """This rank's mailbox handle (64 bytes) for the host-driven bootstrap of the peer-to-peer exchange."""
        buf = (C.c_uint8 * N.COMM_HANDLE_BYTES)()
        N.check(self._lib.rag_comm_p2p_export(self._h, nranks, rank, max_batch, max_k, buf))
        return bytes(buf)

    def comm_p2p_import(self, handles: list[bytes]):
        """All ranks' handles, ordered by rank (what the host all-gathered)."""
        blob = b"".join(handles)
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        N.check(self._lib.rag_comm_p2p_import(self._h, buf))

    def comm_detach(self):
        """Phase 1 of the teardown: close this rank's mappings of the peers' mailboxes (then barrier, then free)."""
        N.check(self._lib.rag_comm_detach(self._h))

    def comm_destroy(self):
        N.check(self._lib.rag_comm_destroy(self._h))


# asynchronous batcher requests in flight: token (the callback's `user` word) -> (result arrays, on_done). ONE module-level
# trampoline serves every request, so no ctypes callback object is ever freed while the library may still call it.
_async_inflight: dict = {}
_async_lock = threading.Lock()
_async_tokens = itertools.count(1)


@N.BATCHER_DONE_FN
def _async_done(user, rc, err):
    with _async_lock:
        out, on_done = _async_inflight.pop(user)
    on_done(out.row(0) if rc == N.OK else N.RagError(rc, (err or b"").decode("utf-8", "replace")))


class Batcher:
    """Micro-batching front end (``rag_batcher_*``): many threads call ``submit`` with one query each; the
    library groups what arrives together into one corpus pass (a batch goes out when the arrivals pause, at the latest
    ``max_wait_us`` after its first request). ctypes releases the GIL during
    the call, so Python request threads really do overlap."""

    def __init__(self, index: VectorIndex, opts: N.HybridOpts, max_batch: int = 1024, max_wait_us: int = 1000):
        self._lib = N.load()
        self.index, self.opts = index, opts
        d = N.BatcherDesc(max_batch, max_wait_us, opts)
        h = C.c_void_p()
        N.check(self._lib.rag_batcher_create(index._h, C.byref(d), C.byref(h)))
        self._h = h

    def submit(self, query, kw_keys=()) -> dict:
        q = np.ascontiguousarray(query, dtype=np.float32).reshape(-1)
        kw = np.ascontiguousarray(kw_keys, dtype=np.uint64)
        out = _alloc_fused(1, max(1, self.opts.vector_top_k + self.opts.keyword_limit + self.opts.fresh_limit),
                           self.opts.vector_top_k)
        N.check(self._lib.rag_batcher_submit(self._h, _ptr(q), _ptr(kw), len(kw), C.byref(out._c)))
        return out.row(0)

    def submit_async(self, query, kw_keys, on_done) -> bool:
        """``rag_batcher_submit_async``: returns at once; ``on_done(result_dict | RagError)`` runs on a batcher worker
        thread when the request's batch has been answered. False = every batch buffer is in flight (``RAG_ERR_BUSY``):
        nothing was queued. One thread can keep thousands of requests in flight this way (the N-API addon's route)."""
        q = np.ascontiguousarray(query, dtype=np.float32).reshape(-1)
        kw = np.ascontiguousarray(kw_keys, dtype=np.uint64)
        out = _alloc_fused(1, max(1, self.opts.vector_top_k + self.opts.keyword_limit + self.opts.fresh_limit),
                           self.opts.vector_top_k)
        token = next(_async_tokens)
        with _async_lock:
            _async_inflight[token] = (out, on_done)     # `out` stays alive until the callback has run
        rc = self._lib.rag_batcher_submit_async(self._h, _ptr(q), _ptr(kw), len(kw), C.byref(out._c), _async_done, token)
        if rc != N.OK:
            with _async_lock:
                _async_inflight.pop(token, None)
            if rc == N.ERR_BUSY:
                return False
            N.check(rc)
        return True

    def stats(self) -> dict:
        a, b, c = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        N.check(self._lib.rag_batcher_stats(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(batches=a.value, queries=b.value, largest_batch=c.value)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.rag_batcher_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

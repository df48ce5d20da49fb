"""ctypes binding of ``libragera.so`` — the C ABI declared in ``include/ragera.h``.

This is the same boundary an N-API addon binds in the reference's Node process
(INTEGRATION.md); Python is used here only because the container has no Node.
There is no CPU fallback: if the shared library is missing ``load()`` raises, and
without an sm_100 GPU ``rag_index_create`` returns ``RAG_ERR_NO_DEVICE``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libragera.so")
CSRC = os.path.join(_HERE, "csrc")

RAGERA_VERSION = 0x00010000
MAX_TOPK, MAX_CANDIDATES, MAX_KEYWORDS, MAX_FRESH = 64, 128, 64, 64
OK, ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_STATE, ERR_UNSUPPORTED, ERR_NCCL, ERR_NO_DEVICE = 0, -1, -2, -3, -4, -5, -6, -7
ERR_TIMEOUT = -8
ERR_BUSY = -9
F32, BF16 = 0, 1
SRC_VECTOR, SRC_KEYWORD, SRC_BOTH, SRC_FRESHNESS = 0, 1, 2, 3
CT_DOCUMENT, CT_MEMORY, CT_CODE = 0, 1, 2
PATH_AUTO, PATH_STREAM, PATH_TENSOR, PATH_EXACT, PATH_SHADOW_STREAM = 0, 1, 2, 3, 4
INDEX_BF16_SHADOW = 1
INDEX_F16_SHADOW = 2
CACHE_META, CACHE_KEYS = 1, 2
SEARCH_NO_ESCALATE = 1
SEARCH_STAT_EPS = 2
PROF_CLASSES = 6
PROF_NAMES = ("stream", "tensor", "merge", "rescore", "fuse", "comm")
COMM_ID_BYTES = 128
COMM_HANDLE_BYTES = 64

class RagError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libragera error {code}: {msg}")
        self.code = code


class GenDesc(C.Structure):
    """rag_gen_desc (include/ragera_gen.h)."""
    _fields_ = [("seed", C.c_uint64), ("query_seed", C.c_uint64), ("meta_seed", C.c_uint64),
                ("total_rows", C.c_uint64), ("n_clusters", C.c_uint32), ("noise", C.c_float),
                ("query_noise", C.c_float), ("dup_period", C.c_uint32), ("memory_rows", C.c_uint64),
                ("now_ms", C.c_int64)]


class IndexDesc(C.Structure):
    _fields_ = [("capacity_rows", C.c_uint64), ("dim", C.c_uint32), ("dtype", C.c_uint32),
                ("device", C.c_int32), ("flags", C.c_uint32), ("id_base", C.c_uint64)]


class RRFConfigC(C.Structure):
    _fields_ = [("k", C.c_double), ("vector_weight", C.c_double), ("keyword_weight", C.c_double),
                ("both_bonus", C.c_double)]


class SearchOpts(C.Structure):
    _fields_ = [("k", C.c_uint32), ("path", C.c_uint32), ("slack", C.c_uint32), ("flags", C.c_uint32),
                ("epsilon", C.c_double)]


class TopkOut(C.Structure):
    _fields_ = [("ids", C.c_void_p), ("scores", C.c_void_p), ("counts", C.c_void_p), ("certified", C.c_void_p)]


class HybridOpts(C.Structure):
    _fields_ = [("vector_top_k", C.c_uint32), ("keyword_limit", C.c_uint32), ("min_vector_score", C.c_double),
                ("rrf", RRFConfigC), ("path", C.c_uint32), ("slack", C.c_uint32), ("flags", C.c_uint32),
                ("fresh_limit", C.c_uint32), ("fresh_weight", C.c_double), ("now_ms", C.c_int64),
                ("time_decay_factor", C.c_double), ("frequency_bonus", C.c_double), ("epsilon", C.c_double)]


class FusedOut(C.Structure):
    _fields_ = [("capacity", C.c_uint32), ("keys", C.c_void_p), ("scores", C.c_void_p), ("source", C.c_void_p),
                ("content_type", C.c_void_p), ("counts", C.c_void_p), ("used_rrf", C.c_void_p),
                ("vec_ids", C.c_void_p), ("vec_scores", C.c_void_p), ("vec_counts", C.c_void_p),
                ("certified", C.c_void_p)]


class Text(C.Structure):
    _fields_ = [("units", C.c_void_p), ("len", C.c_uint32)]


class ProcessOpts(C.Structure):
    _fields_ = [("similarity_threshold", C.c_double), ("min_content_length", C.c_uint32), ("max_results", C.c_uint32),
                ("enable_noise_filter", C.c_uint32), ("enable_rerank", C.c_uint32)]


class ProcessedOut(C.Structure):
    _fields_ = [("capacity", C.c_uint32), ("index", C.c_void_p), ("fusion_score", C.c_void_p), ("deduplicated", C.c_void_p),
                ("source_mask", C.c_void_p), ("n_sources", C.c_void_p), ("count", C.c_uint32)]


class BatcherDesc(C.Structure):
    _fields_ = [("max_batch", C.c_uint32), ("max_wait_us", C.c_uint32), ("opts", HybridOpts)]


# rag_batcher_done_fn(user, rc, err): called on a batcher worker thread (ctypes takes the GIL for the call)
BATCHER_DONE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_char_p)


class MemoryOpts(C.Structure):
    _fields_ = [("limit", C.c_uint32), ("path", C.c_uint32), ("min_relevance", C.c_double), ("now_ms", C.c_int64),
                ("time_decay_factor", C.c_double), ("frequency_bonus", C.c_double), ("similarity_top_k", C.c_uint32),
                ("flags", C.c_uint32)]


class MemoryOut(C.Structure):
    _fields_ = [("ids", C.c_void_p), ("scores", C.c_void_p), ("relevance", C.c_void_p), ("freshness", C.c_void_p),
                ("counts", C.c_void_p)]


# every symbol include/ragera.h declares: name -> (restype, argtypes)
_vp = C.c_void_p
class CacheInfo(C.Structure):
    """rag_cache_info (include/ragera.h): header of a binary sidecar."""
    _fields_ = [("version", C.c_uint32), ("dtype", C.c_uint32), ("dim", C.c_uint32), ("flags", C.c_uint32),
                ("rows", C.c_uint64), ("ids_bytes", C.c_uint64), ("source_size", C.c_uint64),
                ("source_mtime_ns", C.c_int64), ("source_prefix_bytes", C.c_uint64), ("source_prefix_hash", C.c_uint64)]


SYMBOLS = {
    "rag_version": (C.c_int, []),
    "rag_last_error": (C.c_char_p, []),
    "rag_device_count": (C.c_int, []),
    "rag_index_create": (C.c_int, [C.POINTER(IndexDesc), C.POINTER(_vp)]),
    "rag_index_destroy": (None, [_vp]),
    "rag_index_upload": (C.c_int, [_vp, C.c_uint64, C.c_uint64, _vp]),
    "rag_index_generate": (C.c_int, [_vp, C.POINTER(GenDesc), C.c_uint64]),
    "rag_index_set_row_meta": (C.c_int, [_vp, C.c_uint64, C.c_uint64, _vp, _vp, _vp, _vp]),
    "rag_index_set_row_keys": (C.c_int, [_vp, C.c_uint64, C.c_uint64, _vp]),
    "rag_index_read_rows": (C.c_int, [_vp, C.c_uint64, C.c_uint64, _vp]),
    "rag_index_rows": (C.c_uint64, [_vp]),
    "rag_index_read_row_meta": (C.c_int, [_vp, C.c_uint64, C.c_uint64, _vp, _vp, _vp, _vp, _vp]),
    "rag_generate_queries": (C.c_int, [_vp, C.POINTER(GenDesc), C.c_uint64, C.c_uint32, _vp]),
    "rag_index_load_vector_store": (C.c_int, [_vp, C.c_char_p, C.POINTER(C.c_uint64), C.POINTER(_vp), C.POINTER(C.c_uint64)]),
    "rag_parse_vector_store_json": (C.c_int, [C.c_char_p, C.c_uint32, C.c_uint64, _vp, _vp, C.POINTER(C.c_uint64),
                                              C.POINTER(_vp), C.POINTER(C.c_uint64)]),
    "rag_free": (None, [_vp]),
    "rag_parse_vector_store_metadata": (C.c_int, [C.c_char_p, _vp, C.c_uint64, C.c_uint64, _vp, C.POINTER(_vp),
                                                  C.POINTER(C.c_uint64), C.POINTER(C.c_int)]),
    "rag_cache_info_read": (C.c_int, [C.c_char_p, C.POINTER(CacheInfo)]),
    "rag_cache_is_fresh": (C.c_int, [C.c_char_p, C.c_char_p]),
    "rag_cache_refresh_host": (C.c_int, [C.c_char_p, C.c_char_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_int), C.POINTER(C.c_uint64)]),
    "rag_parse_vector_store_json_ex": (C.c_int, [C.c_char_p, C.c_uint32, C.c_uint64, _vp, _vp, C.c_uint64, C.POINTER(C.c_uint64),
                                                 C.POINTER(_vp), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                                 C.POINTER(C.c_int64)]),
    "rag_file_prefix_hash": (C.c_int, [C.c_char_p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_int)]),
    "rag_cache_write_host": (C.c_int, [C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint64, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                       C.c_uint64, C.c_char_p]),
    "rag_cache_read_host": (C.c_int, [C.c_char_p, C.c_uint64, C.c_uint64, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(_vp),
                                      C.POINTER(C.c_uint64)]),
    "rag_index_save_cache": (C.c_int, [_vp, C.c_char_p, _vp, C.c_uint64, C.c_char_p]),
    "rag_index_load_cache": (C.c_int, [_vp, C.c_char_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(_vp),
                                       C.POINTER(C.c_uint64)]),
    "rag_index_open_store": (C.c_int, [_vp, C.c_char_p, C.c_char_p, C.POINTER(C.c_uint64), C.POINTER(_vp),
                                       C.POINTER(C.c_uint64), C.POINTER(C.c_int)]),
    "rag_search": (C.c_int, [_vp, _vp, C.c_uint32, C.POINTER(SearchOpts), C.POINTER(TopkOut)]),
    "rag_hybrid_search": (C.c_int, [_vp, _vp, C.c_uint32, C.POINTER(HybridOpts), _vp, _vp, C.POINTER(FusedOut)]),
    "rag_rrf_fuse": (C.c_int, [_vp, C.c_uint32, C.POINTER(RRFConfigC), _vp, _vp, _vp, C.c_uint32, _vp, _vp,
                               C.c_uint32, C.POINTER(FusedOut)]),
    "rag_memory_retrieve": (C.c_int, [_vp, _vp, C.c_uint32, C.POINTER(MemoryOpts), C.POINTER(MemoryOut)]),
    "rag_freshness_scores": (C.c_int, [_vp, C.c_uint64, _vp, _vp, _vp, C.c_int64, C.c_double, C.c_double, _vp]),
    "rag_stage_batch": (C.c_int, [_vp, _vp, C.c_uint32, _vp, _vp, C.c_uint32]),
    "rag_stage_window": (C.c_int, [_vp, C.c_uint32, C.c_uint32]),
    "rag_hybrid_search_staged": (C.c_int, [_vp, C.c_uint32, C.POINTER(HybridOpts)]),
    "rag_fetch_fused": (C.c_int, [_vp, C.c_uint32, C.POINTER(HybridOpts), C.POINTER(FusedOut)]),
    "rag_sync": (C.c_int, [_vp]),
    "rag_process_results": (C.c_int, [C.POINTER(Text), _vp, _vp, C.c_uint32, Text, C.POINTER(ProcessOpts), C.POINTER(ProcessedOut)]),
    "rag_debug_p2p_next_step": (C.c_uint32, [C.c_uint32]),
    "rag_batcher_create": (C.c_int, [_vp, C.POINTER(BatcherDesc), C.POINTER(_vp)]),
    "rag_batcher_submit": (C.c_int, [_vp, _vp, _vp, C.c_uint32, C.POINTER(FusedOut)]),
    "rag_batcher_submit_async": (C.c_int, [_vp, _vp, _vp, C.c_uint32, C.POINTER(FusedOut), BATCHER_DONE_FN, _vp]),
    "rag_batcher_stats": (C.c_int, [_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "rag_batcher_destroy": (None, [_vp]),
    "rag_debug_tensor_scores": (C.c_int, [_vp, _vp, C.c_uint32, _vp]),
    "rag_debug_tensor_candidates": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, _vp, _vp]),
    "rag_timer_start": (C.c_int, [_vp]),
    "rag_timer_stop": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "rag_index_row_residual": (C.c_double, [_vp]),
    "rag_launch_count": (C.c_uint64, [_vp]),
    "rag_certified_totals": (C.c_int, [_vp, _vp, _vp]),
    "rag_profile_enable": (C.c_int, [_vp, C.c_int]),
    "rag_profile_read": (C.c_int, [_vp, C.POINTER(C.c_float), C.POINTER(C.c_uint32)]),
    "rag_host_alloc": (_vp, [C.c_uint64]),
    "rag_host_free": (None, [_vp]),
    "rag_comm_p2p_export": (C.c_int, [_vp, C.c_int, C.c_int, C.c_uint32, C.c_uint32, _vp]),
    "rag_comm_p2p_import": (C.c_int, [_vp, _vp]),
    "rag_comm_detach": (C.c_int, [_vp]),
    "rag_comm_unique_id": (C.c_int, [_vp]),
    "rag_comm_init": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "rag_comm_destroy": (C.c_int, [_vp]),
}


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libragera.so for sm_100a with the committed Makefile (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    srcs += [os.path.join(_HERE, "..", "include", f) for f in ("ragera.h", "ragera_gen.h")]
    stale = (not os.path.exists(LIB_PATH)) or any(
        os.path.getmtime(p) > os.path.getmtime(LIB_PATH) for p in srcs if os.path.exists(p))
    if force or stale:
        cmd = ["make", "-C", CSRC, "-j", str(min(16, os.cpu_count() or 4))] + (["-B"] if force else [])
        r = subprocess.run(cmd, capture_output=not verbose, text=True)
        if r.returncode != 0:
            raise RuntimeError("building libragera.so failed:\n" + (r.stdout or "") + (r.stderr or ""))
    return LIB_PATH


_lib = None


def load() -> C.CDLL:
    """Load libragera.so and type every entry point. Raises if it is not built — there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C rag_era_b200/csrc`). rag_era_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        if lib.rag_version() != RAGERA_VERSION:
            raise RuntimeError(f"libragera.so version {lib.rag_version():#x} != binding {RAGERA_VERSION:#x}")
        _lib = lib
    return _lib


ON_ROWS = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_float))


def parse_vector_store_json(path: str, dim: int, slab_rows: int = 4096, resume_offset: int | None = None, keep_rows: bool = True):
    """Host-only parse of a llamaindex vector_store.json → (ids, rows float32 [n, dim]). No GPU needed.
    With ``resume_offset`` (0 = from the start, else the end offset an earlier parse returned) the result is
    (ids, rows, end_offset): only the embeddings after that offset are parsed."""
    import numpy as np

    lib = load()
    chunks = []

    def on_rows(_user, first, n, ptr):
        if keep_rows:
            chunks.append(np.ctypeslib.as_array(ptr, shape=(n * dim,)).reshape(n, dim).copy())
        return OK

    cb = ON_ROWS(on_rows)
    rows, blob, nbytes, end = C.c_uint64(0), C.c_void_p(), C.c_uint64(0), C.c_uint64(0)
    if resume_offset is None:
        check(lib.rag_parse_vector_store_json(path.encode(), dim, slab_rows, C.cast(cb, C.c_void_p), None, C.byref(rows),
                                              C.byref(blob), C.byref(nbytes)))
    else:
        check(lib.rag_parse_vector_store_json_ex(path.encode(), dim, slab_rows, C.cast(cb, C.c_void_p), None, resume_offset,
                                                 C.byref(rows), C.byref(blob), C.byref(nbytes), C.byref(end), None, None))
    ids = C.string_at(blob, nbytes.value).decode("utf-8").split("\0")[:-1] if nbytes.value else []
    lib.rag_free(blob)
    X = np.vstack(chunks) if chunks else np.zeros((0, dim), np.float32)
    return (ids, X) if resume_offset is None else (ids, X, int(end.value))


def _ids_blob(ids) -> bytes:
    return b"".join(i.encode("utf-8") + b"\0" for i in ids)


def _split_blob(blob, nbytes) -> list:
    out = C.string_at(blob, nbytes).decode("utf-8").split("\0")[:-1] if nbytes else []
    load().rag_free(blob)
    return out


def parse_vector_store_metadata(path: str, ids: list):
    """Host-only second pass: ``metadataDict`` → (content_type uint8 [rows], memory ids [rows], found)."""
    import numpy as np

    blob = _ids_blob(ids)
    ct = np.zeros(len(ids), np.uint8)
    mem, nbytes, found = C.c_void_p(), C.c_uint64(0), C.c_int(0)
    check(load().rag_parse_vector_store_metadata(path.encode(), blob, len(blob), len(ids), ct.ctypes.data_as(C.c_void_p),
                                                 C.byref(mem), C.byref(nbytes), C.byref(found)))
    return ct, _split_blob(mem, nbytes.value), bool(found.value)


def cache_info(path: str) -> CacheInfo:
    info = CacheInfo()
    check(load().rag_cache_info_read(path.encode(), C.byref(info)))
    return info


def cache_refresh(source_json: str, dtype: int, dim: int, cache_path: str | None = None) -> tuple[int, int]:
    """rag_cache_refresh_host: bring the sidecar up to date without a GPU. Returns (route, rows): route 1 = fresh,
    2 = extended in place with the embeddings appended to the JSON, 0 = full parse + rewrite."""
    route, rows = C.c_int(0), C.c_uint64(0)
    check(load().rag_cache_refresh_host(cache_path.encode() if cache_path else None, source_json.encode(), dtype, dim,
                                        C.byref(route), C.byref(rows)))
    return int(route.value), int(rows.value)


def cache_is_fresh(cache_path: str, source_json: str) -> bool:
    return bool(load().rag_cache_is_fresh(cache_path.encode(), source_json.encode()))


def cache_write_host(path: str, rows, ids=None, meta=None, keys=None, source_json=None, dtype=F32):
    """Host-only writer of the binary sidecar. rows: float32 [n, dim] (or uint16 bf16 bits with dtype=BF16);
    meta: (content_type u8, confidence f64, access_count i32, last_access_ms i64) or None."""
    import numpy as np

    rows = np.ascontiguousarray(rows, dtype=np.uint16 if dtype == BF16 else np.float32)
    n, dim = rows.shape
    ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    m = [None] * 4 if meta is None else [np.ascontiguousarray(a, dtype=t) for a, t in
                                         zip(meta, (np.uint8, np.float64, np.int32, np.int64))]
    k = None if keys is None else np.ascontiguousarray(keys, dtype=np.uint64)
    blob = _ids_blob(ids) if ids is not None else None
    check(load().rag_cache_write_host(path.encode(), dtype, dim, n, ptr(rows), *[ptr(a) for a in m], ptr(k), blob,
                                      len(blob) if blob else 0, source_json.encode() if source_json else None))


def cache_read_host(path: str, first_row: int = 0, nrows: int | None = None) -> dict:
    """Host-only reader (every checksum verified) → dict(rows, ids, content_type, confidence, access_count,
    last_access_ms, keys); absent sections are None."""
    import numpy as np

    info = cache_info(path)
    n = info.rows - first_row if nrows is None else nrows
    rows = np.empty((n, info.dim), np.uint16 if info.dtype == BF16 else np.float32)
    has_meta, has_keys = bool(info.flags & CACHE_META), bool(info.flags & CACHE_KEYS)
    m = [np.empty(n, t) if has_meta else None for t in (np.uint8, np.float64, np.int32, np.int64)]
    k = np.empty(n, np.uint64) if has_keys else None
    ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    blob, nbytes = C.c_void_p(), C.c_uint64(0)
    check(load().rag_cache_read_host(path.encode(), first_row, n, ptr(rows), *[ptr(a) for a in m], ptr(k), C.byref(blob),
                                     C.byref(nbytes)))
    return dict(rows=rows, ids=_split_blob(blob, nbytes.value), content_type=m[0], confidence=m[1], access_count=m[2],
                last_access_ms=m[3], keys=k, info=info)


def check(rc: int) -> None:
    if rc != OK:
        raise RagError(rc, load().rag_last_error().decode("utf-8", "replace"))

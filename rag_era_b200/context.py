"""Host-side mirror of ``src/lib/context/rag/dedup-filter.ts`` (``processResults``) over the C ABI
(``rag_process_results``, rag_era_b200/csrc/postfilter.cu) — SURVEY §8f N3. No GPU involved."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _native as N

SOURCE_ID = {"vector": 0, "keyword": 1, "graph": 2, "hybrid": 3, "both": 3}


@dataclass
class FusedResult:
    """FusedResult — src/lib/context/types.ts:79-83 (plus the index of the input result it came from)."""
    index: int
    content: str
    score: float
    fusionScore: float
    sources: int
    deduplicated: bool


def _utf16(s: str) -> np.ndarray:
    return np.frombuffer(s.encode("utf-16-le", "surrogatepass"), dtype=np.uint16).copy()


def process_results(results: Sequence, query: str, dedup_config: Optional[dict] = None, enable_noise_filter: bool = True,
                    enable_rerank: bool = True) -> list[FusedResult]:
    """processResults(results, query, options) — dedup-filter.ts:193-247. ``results`` have .content, .score, .source."""
    lib = N.load()
    n = len(results)
    bufs = [_utf16(r.content) for r in results]
    texts = (N.Text * max(n, 1))(*[N.Text(b.ctypes.data, len(b)) for b in bufs])
    scores = np.array([float(r.score) for r in results], dtype=np.float64)
    sources = np.array([SOURCE_ID.get(getattr(r, "source", "vector"), 0) for r in results], dtype=np.uint8)
    qb = _utf16(query)
    cfg = dict(similarityThreshold=0.85, minContentLength=20, maxResults=10)
    cfg.update(dedup_config or {})
    opts = N.ProcessOpts(cfg["similarityThreshold"], cfg["minContentLength"], cfg["maxResults"], int(enable_noise_filter), int(enable_rerank))
    cap = max(n, 1)
    idx = np.zeros(cap, np.uint32); fs = np.zeros(cap, np.float64); dd = np.zeros(cap, np.uint8)
    sm = np.zeros(cap, np.uint32); ns = np.zeros(cap, np.uint32)
    out = N.ProcessedOut(cap, idx.ctypes.data, fs.ctypes.data, dd.ctypes.data, sm.ctypes.data, ns.ctypes.data, 0)
    N.check(lib.rag_process_results(texts, scores.ctypes.data, sources.ctypes.data, n, N.Text(qb.ctypes.data, len(qb)),
                                    C.byref(opts), C.byref(out)))
    return [FusedResult(int(idx[i]), results[int(idx[i])].content, float(scores[int(idx[i])]), float(fs[i]), int(ns[i]), bool(dd[i]))
            for i in range(out.count)]

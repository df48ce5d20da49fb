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


# ---- the callers of the hot path (north_star: "behind ContextEngine.buildContext and the
#      search_knowledge (top-3) and deep_search (top-8) tools"). Host glue only: every score comes from the device. ----
import math


@dataclass
class RetrievalDecision:
    """RetrievalDecision — src/lib/context/rag/retrieval-decision.ts (fields read by calculateRetrievalCount)."""
    shouldRetrieve: bool = True
    reason: str = "默认混合检索"
    queryType: str = "hybrid"        # 'semantic' | 'keyword' | 'graph' | 'hybrid'
    estimatedResults: int = 8
    priority: str = "medium"         # 'high' | 'medium' | 'low'


def calculate_retrieval_count(decision: RetrievalDecision, max_token_budget: int = 2000, average_chunk_tokens: int = 150) -> dict:
    """calculateRetrievalCount — retrieval-decision.ts:144-195."""
    max_chunks = math.floor(max_token_budget / average_chunk_tokens)
    mult = 1.5 if decision.priority == "high" else 1.0 if decision.priority == "medium" else 0.7
    base = math.floor(max_chunks * mult)
    if decision.queryType == "semantic":
        return dict(vectorTopK=base, keywordLimit=0, graphLimit=0)
    if decision.queryType == "keyword":
        return dict(vectorTopK=2, keywordLimit=base, graphLimit=0)
    if decision.queryType == "graph":
        return dict(vectorTopK=3, keywordLimit=0, graphLimit=base)
    return dict(vectorTopK=math.ceil(base * 0.6), keywordLimit=math.ceil(base * 0.4), graphLimit=0)


@dataclass
class SearchResult:
    """SearchResult of the context engine — src/lib/context/types.ts (id, content, documentName, score, source)."""
    id: str
    content: str
    documentName: str
    score: float
    source: str


_SOURCE_MAP = {"vector": "vector", "keyword": "keyword", "both": "hybrid", "graph": "graph", "hybrid": "hybrid"}


def get_unified_results(index, knowledge_base_id: str, query, decision: Optional[RetrievalDecision] = None, *,
                        keyword_service=None, now_ms: int = 0, path: int = N.PATH_AUTO) -> dict:
    """ContextEngine.getUnifiedResults — src/lib/context/engine.ts:225-299: one hybridSearch with
    (vectorTopK = c + 10, keywordLimit = c', minVectorScore = 0.4), split by contentType, memories carry the RRF score
    with the constants of :255-266, documents go through processResults (:289); at most 10 memories (:292).
    Any failure returns empty lists like the reference's catch (:295-298)."""
    from .hybrid_search import hybrid_search
    from .memory import ScoredMemory

    try:
        counts = calculate_retrieval_count(decision or RetrievalDecision())
        results = hybrid_search(index, knowledge_base_id, query,
                                dict(vectorTopK=counts["vectorTopK"] + 10, keywordLimit=counts["keywordLimit"], minVectorScore=0.4),
                                keyword_service=keyword_service, path=path)
        memories, documents = [], []
        for r in results:
            if r.contentType == "memory":
                md = r.metadata or {}
                memories.append(ScoredMemory(id=md.get("memoryId") or r.id, knowledgeBaseId=knowledge_base_id, content=r.content,
                                             type=md.get("memoryType") or "context", confidence=0.8, accessCount=0,
                                             lastAccessedAt=now_ms, score=r.score, relevanceScore=r.score, freshnessScore=0.5))
            else:
                documents.append(SearchResult(r.id, r.content, r.documentName, r.score, _SOURCE_MAP.get(r.source) or "hybrid"))
        text_query = query if isinstance(query, str) else ""
        processed = process_results(documents, text_query)
        return dict(memories=memories[:10], documents=processed, raw_documents=documents)
    except N.RagError:
        return dict(memories=[], documents=[], raw_documents=[])


@dataclass
class ToolContext:
    """ToolContext — src/lib/llm/tools/types.ts (the fields the search tools touch)."""
    index: object
    knowledgeBaseId: str
    keyword_service: object = None
    toolCalls: list = field(default_factory=list)
    searchResults: list = field(default_factory=list)


def _search_tool(ctx: ToolContext, name: str, query, top_k: int, show: int, path: int) -> str:
    from .hybrid_search import format_search_results, hybrid_search, js_substring

    results = hybrid_search(ctx.index, ctx.knowledgeBaseId, query, dict(vectorTopK=top_k, keywordLimit=top_k),
                            keyword_service=ctx.keyword_service, path=path)
    if not results:
        ctx.toolCalls.append(dict(tool=name, input=query, output="未找到相关内容"))
        return "未找到相关内容"
    formatted = format_search_results(results, show)
    ctx.toolCalls.append(dict(tool=name, input=query, output=js_substring(formatted, 0, 200)))
    if not ctx.searchResults:
        ctx.searchResults.extend(results)
    return formatted


def search_knowledge(ctx: ToolContext, query, path: int = N.PATH_AUTO) -> str:
    """search_knowledge — src/lib/llm/tools/search-tools.ts:12-51: hybridSearch(vectorTopK 5, keywordLimit 5), show top 3."""
    return _search_tool(ctx, "search_knowledge", query, 5, 3, path)


def deep_search(ctx: ToolContext, query, path: int = N.PATH_AUTO) -> str:
    """deep_search — src/lib/llm/tools/search-tools.ts:56-95: hybridSearch(vectorTopK 10, keywordLimit 10), show top 8."""
    return _search_tool(ctx, "deep_search", query, 10, 8, path)

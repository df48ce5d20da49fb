"""rag_era_b200 — B200-native (sm_100a) implementation of gong9/rag-era's retrieval hot path:
dense cosine scoring → per-query top-k → min-cosine filter → Reciprocal Rank Fusion
(+ memory freshness), behind the reference's ``hybridSearch`` signature.

The product is ``libragera.so`` (CUDA kernels + C ABI, ``include/ragera.h``); this package is
the host-side mirror of the reference's TypeScript interface used by tests and benchmarks.
There is no CPU implementation: importing works anywhere, calling needs a B200.
"""
from . import _native
from ._native import RagError
from .index import Batcher, VectorIndex, RRFConfig, TopK, Fused, hybrid_opts
from .hybrid_search import (PRESET_CONFIGS, HybridSearchResult, KeywordHit, KnowledgeIndex, Node, NodeWithScore, Retriever,
                            format_search_results,
                            get_preset_config, get_source_stats, hybrid_search, reciprocal_rank_fusion)
from .memory import (Memory, MemoryStore, ScoredMemory, batch_calculate_freshness, calculate_freshness_score,
                     sort_by_freshness)
from .sharded import create_sharded_index, open_sharded_cache, shard_range
from .context import (FusedResult, RetrievalDecision, SearchResult, ToolContext, calculate_retrieval_count, deep_search,
                      get_unified_results, process_results, search_knowledge)

__all__ = [
    "RagError", "Batcher", "VectorIndex", "RRFConfig", "TopK", "Fused", "hybrid_opts", "PRESET_CONFIGS", "HybridSearchResult",
    "KeywordHit", "KnowledgeIndex", "Node", "NodeWithScore", "Retriever", "format_search_results", "get_preset_config", "get_source_stats",
    "hybrid_search", "reciprocal_rank_fusion", "Memory", "MemoryStore", "ScoredMemory", "batch_calculate_freshness",
    "calculate_freshness_score", "sort_by_freshness", "create_sharded_index", "open_sharded_cache", "shard_range", "FusedResult", "process_results", "RetrievalDecision",
    "SearchResult", "ToolContext", "calculate_retrieval_count", "deep_search", "get_unified_results", "search_knowledge",
]

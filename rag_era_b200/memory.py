"""Host-side mirror of ``src/lib/memory/{freshness,store}.ts`` over the C ABI.

    calculate_freshness_score(memory, now, config)   calculateFreshnessScore  freshness.ts:37-56
    MemoryStore.retrieve(query, limit, min_relevance) MemoryStore.retrieve     store.ts:102-180

The exp/log arithmetic and the 0.7/0.3 blend run on the GPU (``rag_freshness_scores``,
``rag_memory_retrieve``); the host holds the Memory records the reference keeps in SQLite
(``prisma/schema.prisma:87-106``).
"""
from __future__ import annotations

from dataclasses import dataclass, replace
from typing import Optional, Sequence

import numpy as np

from . import _native as N
from .hybrid_search import KnowledgeIndex, Node

DEFAULT_CONFIG = dict(timeDecayFactor=0.05, frequencyBonus=0.1)  # freshness.ts:20-23


@dataclass
class Memory:
    """Memory — src/lib/memory/types.ts:22-32 (the columns the hot path reads)."""
    id: str
    knowledgeBaseId: str
    content: str
    confidence: float
    accessCount: int
    lastAccessedAt: int  # epoch milliseconds (Date.getTime())
    type: str = "context"  # MemoryType (types.ts:10-17); carried, never read by the hot path


@dataclass
class ScoredMemory(Memory):
    """ScoredMemory — src/lib/memory/types.ts:37-41."""
    score: float = 0.0
    relevanceScore: float = 0.0
    freshnessScore: float = 0.0


def calculate_freshness_score(memory: Memory, now_ms: int, config: Optional[dict] = None, *, store) -> float:
    """calculateFreshnessScore — freshness.ts:37-56; ``store`` is any VectorIndex (supplies the stream)."""
    cfg = config or DEFAULT_CONFIG
    return float(store.freshness_scores([memory.confidence], [memory.accessCount], [memory.lastAccessedAt], now_ms,
                                        cfg["timeDecayFactor"], cfg["frequencyBonus"])[0])


def batch_calculate_freshness(memories: Sequence[Memory], now_ms: int, *, store):
    """batchCalculateFreshness — freshness.ts:61-69."""
    if not memories:
        return []
    sc = store.freshness_scores([m.confidence for m in memories], [m.accessCount for m in memories],
                                [m.lastAccessedAt for m in memories], now_ms)
    return [dict(memory=m, freshnessScore=float(s)) for m, s in zip(memories, sc)]


def sort_by_freshness(memories: Sequence[Memory], now_ms: int, *, store) -> list:
    """sortByFreshness — freshness.ts:74-83: freshness from the device, V8's stable sort (ties keep their order)."""
    scored = batch_calculate_freshness(memories, now_ms, store=store)
    return [s["memory"] for s in sorted(scored, key=lambda s: -s["freshnessScore"])]


class MemoryStore:
    """MemoryStore of src/lib/memory/store.ts bound to one knowledge base's unified index."""

    def __init__(self, knowledge_base_id: str, index: KnowledgeIndex):
        self.knowledge_base_id = knowledge_base_id
        self.index = index
        self._db: dict[str, Memory] = {}        # prisma.memory
        self._row_memory: dict[int, str] = {}   # row → memoryId (node.metadata.memoryId)

    def store(self, memory: Memory, embedding) -> None:
        """store — store.ts:43-77: the memory becomes one more row of the SAME index (index.insert)."""
        node = Node(id_=f"memory_{memory.id}", text=memory.content,
                    metadata=dict(type="memory", memoryId=memory.id, knowledgeBaseId=memory.knowledgeBaseId))
        row0 = self.index.insert_nodes([node], np.asarray(embedding, dtype=np.float32)[None, :])
        self.index.store.set_row_meta(row0, content_type=[N.CT_MEMORY], confidence=[memory.confidence],
                                      access_count=[memory.accessCount], last_access_ms=[memory.lastAccessedAt])
        self._db[memory.id] = memory
        self._row_memory[row0] = memory.id

    def retrieve(self, query, limit: int = 10, min_relevance: float = 0.5, now_ms: int = 0) -> list[ScoredMemory]:
        """retrieve — store.ts:102-180. Errors degrade to [] exactly like the reference's catch (:176-179).

        The device does top-2·limit → memory rows → cos ≥ minRelevance → 0.7·cos + 0.3·fresh → stable sort. Rows whose DB
        record is gone (``delete`` leaves the vector node behind, :240-251) or belongs to another knowledge base are
        dropped HERE, after the sort and before the slice — the same set and order as ``if (dbMemory)`` (:153) followed
        by ``sort`` and ``slice(0, limit)`` — so the device is asked for all 2·limit blended candidates."""
        try:
            q = self.index.embed(query)
            r = self.index.store.memory_retrieve(q, 2 * limit, min_relevance, now_ms, similarity_top_k=2 * limit)
            n = int(r["counts"][0])
            out = []
            for i in range(n):
                row = int(r["ids"][0, i]) - self.index.store.id_base
                mem = self._db.get(self._row_memory.get(row, ""))
                if mem is None or mem.knowledgeBaseId != self.knowledge_base_id:   # :119-122, :153
                    continue
                out.append(ScoredMemory(**vars(mem), score=float(r["scores"][0, i]),
                                        relevanceScore=float(r["relevance"][0, i]),
                                        freshnessScore=float(r["freshness"][0, i])))
            return out[:limit]
        except N.RagError:
            return []

    def _push_meta(self, memory_id: str) -> None:
        m = self._db[memory_id]
        for row, mid in self._row_memory.items():
            if mid == memory_id:
                self.index.store.set_row_meta(row, content_type=[N.CT_MEMORY], confidence=[m.confidence],
                                              access_count=[m.accessCount], last_access_ms=[m.lastAccessedAt])

    def touch(self, memory_id: str, now_ms: int) -> None:
        """touch — store.ts:207-215: accessCount += 1, lastAccessedAt = now; the device copy of the row follows, so the
        next retrieve computes freshness from the new values."""
        m = self._db[memory_id]
        m.accessCount += 1
        m.lastAccessedAt = now_ms
        self._push_meta(memory_id)

    def touch_many(self, memory_ids: Sequence[str], now_ms: int) -> None:
        """touchMany — store.ts:220-235."""
        for mid in memory_ids:
            self.touch(mid, now_ms)

    def delete(self, memory_id: str) -> None:
        """delete — store.ts:240-251: only the DB record goes; the vector node stays in the index (it keeps showing up
        in hybridSearch as a memory hit, exactly as in the reference) and ``retrieve`` filters it out."""
        del self._db[memory_id]

    def get_all(self) -> list[Memory]:
        """getAll — store.ts:254-260."""
        return [m for m in self._db.values() if m.knowledgeBaseId == self.knowledge_base_id]

    def count(self) -> int:
        """count — store.ts:265-269."""
        return len(self.get_all())

    def has_similar(self, content_embedding, threshold: float = 0.9, now_ms: int = 0) -> bool:
        """hasSimilar — store.ts:274-285: retrieve(content, 1) and compare relevance."""
        r = self.retrieve(content_embedding, 1, 0.5, now_ms)
        return bool(r) and r[0].relevanceScore >= threshold

"""Host-side mirror of the reference's ``src/lib/hybrid-search.ts`` over the C ABI.

Same names, argument meaning and error behaviour as the TypeScript module so the parity
tests read like tests of the reference:

    hybrid_search(index, knowledge_base_id, query, options)   hybridSearch        :275-355
    reciprocal_rank_fusion(vector, keyword, config)          reciprocalRankFusion :129-208
    get_preset_config / format_search_results / get_source_stats   :110, :364, :378

What runs where: the dense scoring, top-k, min-cosine filter, RRF arithmetic and the final
stable sort run on the GPU through ``rag_hybrid_search`` / ``rag_rrf_fuse``. The host only
does what the JS adapter would do: embed the query (injected callable — a network call in
the reference), fetch the keyword hits (injected service — Meilisearch in the reference),
turn ``content.substring(0, 100)`` keys into integers, and re-attach strings to the fused
integer keys. No score is computed on the host.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Callable, Optional, Sequence

import numpy as np

from . import _native as N
from .index import RRFConfig, VectorIndex, hybrid_opts

SOURCE = {N.SRC_VECTOR: "vector", N.SRC_KEYWORD: "keyword", N.SRC_BOTH: "both", N.SRC_FRESHNESS: "freshness"}
CONTENT_TYPE = {N.CT_DOCUMENT: "document", N.CT_MEMORY: "memory", N.CT_CODE: "code"}
CONTENT_TYPE_ID = {v: k for k, v in CONTENT_TYPE.items()}

# PRESET_CONFIGS — src/lib/hybrid-search.ts:77-105
PRESET_CONFIGS = {
    "document": dict(rrf=RRFConfig(60.0, 1.0, 1.0, 0.1), vectorTopK=8, keywordLimit=8, minVectorScore=0.3),
    "code": dict(rrf=RRFConfig(40.0, 1.0, 1.3, 0.15), vectorTopK=6, keywordLimit=5, minVectorScore=0.25),
}


def get_preset_config(preset: str):
    """getPresetConfig — src/lib/hybrid-search.ts:110-112."""
    return PRESET_CONFIGS[preset]


def js_substring(s: str, start: int, end: int) -> str:
    """``String.prototype.substring`` counts UTF-16 code units, not code points (:149,:171)."""
    u = s.encode("utf-16-le", "surrogatepass")
    return u[2 * start:2 * end].decode("utf-16-le", "surrogatepass")


@dataclass
class HybridSearchResult:
    """HybridSearchResult — src/lib/hybrid-search.ts:18-27."""
    id: str
    documentName: str
    content: str
    score: float
    source: str
    contentType: str
    documentId: Optional[str] = None
    metadata: Optional[dict] = None


@dataclass
class KeywordHit:
    """SearchResult of src/lib/meilisearch.ts:226-236 (``score`` is ignored by RRF)."""
    id: str
    documentId: str
    documentName: str
    content: str
    score: float = 0.0


@dataclass
class Node:
    """What the retriever hands back per hit: node id, text, metadata (hybrid-search.ts:226-246)."""
    id_: str
    text: str
    metadata: dict = field(default_factory=dict)


class KeyInterner:
    """Opaque integer fusion keys for ``content.substring(0,100)`` (SURVEY N-key)."""

    def __init__(self):
        self._ids: dict[str, int] = {}
        self._strs: list[str] = []

    def key(self, content: str) -> int:
        k = js_substring(content, 0, 100)
        i = self._ids.get(k)
        if i is None:
            i = len(self._strs)
            self._ids[k] = i
            self._strs.append(k)
        return i

    def string(self, i: int) -> str:
        return self._strs[i]


class KnowledgeIndex:
    """VectorStoreIndex stand-in: device matrix + the host-side node table.

    ``embed_model`` plays ``Settings.embedModel.getQueryEmbedding`` (a network call in the
    reference, ``src/lib/llm/config.ts:63-67``): any callable ``str -> vector``.
    """

    def __init__(self, dim: int, capacity_rows: int, embed_model: Optional[Callable[[str], Sequence[float]]] = None,
                 dtype: int = N.F32, device: int = 0, bf16_shadow: bool = False, shadow: str | None = None):
        self.store = VectorIndex(dim, capacity_rows, dtype=dtype, device=device, bf16_shadow=bf16_shadow, shadow=shadow)
        self.embed_model = embed_model
        self.nodes: list[Node] = []
        self.keys = KeyInterner()

    def close(self):
        self.store.close()

    def insert_nodes(self, nodes: Sequence[Node], embeddings: np.ndarray, is_codebase: bool = False):
        """VectorStoreIndex.fromDocuments / index.insert: rows are appended in order."""
        if len(nodes) != len(embeddings):
            raise ValueError("one embedding per node")
        row0 = self.store.upload(embeddings)
        ctype = np.array([CONTENT_TYPE_ID[classify_content_type(n.metadata, is_codebase)] for n in nodes], dtype=np.uint8)
        self.store.set_row_meta(row0, content_type=ctype, confidence=np.zeros(len(nodes)),
                                access_count=np.zeros(len(nodes), np.int32), last_access_ms=np.zeros(len(nodes), np.int64))
        self.store.set_row_keys(row0, [self.keys.key(n.text) for n in nodes])
        self.nodes.extend(nodes)
        return row0

    def as_retriever(self, similarity_top_k: int = 2, path: int = N.PATH_AUTO) -> "Retriever":
        """index.asRetriever({ similarityTopK }) — llamaindex's default similarityTopK is 2."""
        return Retriever(self, similarity_top_k, path)

    def embed(self, query) -> np.ndarray:
        if isinstance(query, str):
            if self.embed_model is None:
                raise ValueError("a string query needs an embed_model")
            query = self.embed_model(query)
        return np.ascontiguousarray(query, dtype=np.float32)


@dataclass
class NodeWithScore:
    """What ``retriever.retrieve`` returns per hit (llamaindex NodeWithScore): the node and its cosine."""
    node: Node
    score: float


class Retriever:
    """``index.asRetriever({ similarityTopK })`` (hybrid-search.ts:223, memory/store.ts:111-113, summarize-tool.ts:46):
    ``retrieve(query)`` embeds the query and returns the top-k nodes in rank order with exact fp64 cosines —
    SimpleVectorStore.query → getTopKEmbeddings, run by ``rag_search`` on the device."""

    def __init__(self, index: "KnowledgeIndex", similarity_top_k: int = 2, path: int = N.PATH_AUTO):
        self.index, self.similarity_top_k, self.path = index, similarity_top_k, path

    def retrieve(self, query) -> list[NodeWithScore]:
        q = self.index.embed(query)
        ids, scores = self.index.store.query(q, self.similarity_top_k, path=self.path).row(0)
        base = self.index.store.id_base
        return [NodeWithScore(self.index.nodes[int(i) - base], float(s)) for i, s in zip(ids, scores)]


def classify_content_type(metadata: dict, is_codebase: bool) -> str:
    """hybrid-search.ts:229-234. ``metadata.language !== undefined``: a present key counts even when its value is
    null or '' (JSON cannot hold undefined, so presence is the test)."""
    if (metadata or {}).get("type") == "memory":
        return "memory"
    if is_codebase or "language" in (metadata or {}):
        return "code"
    return "document"


def _document_name(metadata: dict, is_memory: bool) -> str:
    """hybrid-search.ts:238-240."""
    if is_memory:
        return "用户记忆"
    return metadata.get("documentName") or metadata.get("relativePath") or metadata.get("filePath") or "未知文档"


def reciprocal_rank_fusion(vector_results: Sequence[Any], keyword_results: Sequence[Any],
                           config: RRFConfig = PRESET_CONFIGS["document"]["rrf"],
                           *, store: VectorIndex) -> list[HybridSearchResult]:
    """reciprocalRankFusion — src/lib/hybrid-search.ts:129-208, scores computed by ``rag_rrf_fuse``.

    ``vector_results`` / ``keyword_results`` are objects with the VectorResult / SearchResult
    fields (``content``, ``documentName``, ``contentType``, ``metadata``, ``documentId``).
    """
    keys = KeyInterner()
    vk = [keys.key(r.content) for r in vector_results]
    kk = [keys.key(r.content) for r in keyword_results]
    vt = [CONTENT_TYPE_ID[getattr(r, "contentType", "document")] for r in vector_results]
    fused = store.rrf_fuse([vk], [kk], config, vec_ctypes=[vt]).row(0)
    first: dict[int, tuple[str, Any]] = {}
    for r, k in zip(vector_results, vk):
        first.setdefault(k, ("vector", r))
    for r, k in zip(keyword_results, kk):
        first.setdefault(k, ("keyword", r))
    out = []
    for key, score, src, ct in zip(fused["keys"], fused["scores"], fused["source"], fused["ctype"]):
        origin, r = first[int(key)]
        out.append(HybridSearchResult(
            id=keys.string(int(key)), documentName=r.documentName, content=r.content, score=float(score),
            source=SOURCE[int(src)], contentType=CONTENT_TYPE[int(ct)],
            documentId=getattr(r, "documentId", None) if origin == "keyword" else None,
            metadata=getattr(r, "metadata", None) if origin == "vector" else None))
    return out


def hybrid_search(index: KnowledgeIndex, knowledge_base_id: str, query, options: Optional[dict] = None,
                  *, keyword_service=None, path: int = N.PATH_AUTO) -> list[HybridSearchResult]:
    """hybridSearch — src/lib/hybrid-search.ts:275-355.

    ``options`` carries the HybridSearchOptions keys (vectorTopK, keywordLimit, useKeyword,
    minVectorScore, preset, rrfConfig); ``None`` values fall back to the preset like ``??``
    (so ``keywordLimit: 0`` is respected). ``keyword_service`` is the meilisearchService
    stand-in: ``is_available()`` and ``search(kb, query, limit) -> [KeywordHit]``.
    """
    options = options or {}
    preset = options.get("preset") or "document"
    pc = PRESET_CONFIGS[preset]
    # :296 — a per-call flag: every non-memory VECTOR hit of a codebase search is 'code' (:229-234); rows inserted with
    # is_codebase or carrying metadata.language already read as code from the device
    is_codebase = preset == "code" or knowledge_base_id.startswith("codebase_")
    ctype_of = lambda ct: "code" if is_codebase and CONTENT_TYPE[int(ct)] == "document" else CONTENT_TYPE[int(ct)]
    opt = lambda name: options[name] if options.get(name) is not None else pc[name]
    vector_top_k, keyword_limit, min_vector_score = opt("vectorTopK"), opt("keywordLimit"), opt("minVectorScore")
    use_keyword = options["useKeyword"] if options.get("useKeyword") is not None else True
    rrf = RRFConfig(pc["rrf"].k, pc["rrf"].vector_weight, pc["rrf"].keyword_weight, pc["rrf"].both_bonus)
    for js, py in (("k", "k"), ("vectorWeight", "vector_weight"), ("keywordWeight", "keyword_weight"),
                   ("bothBonus", "both_bonus")):
        if (options.get("rrfConfig") or {}).get(js) is not None:
            setattr(rrf, py, float(options["rrfConfig"][js]))

    # 3. keyword hits first on the host (order of the network calls does not matter to the result)
    keyword_results: list[KeywordHit] = []
    if use_keyword and keyword_service is not None and keyword_service.is_available():
        keyword_results = list(keyword_service.search(knowledge_base_id, query, keyword_limit))
    kw_keys = [index.keys.key(h.content) for h in keyword_results]

    # 1.+2.+4. on the device: top-k, min-cosine filter, RRF (or the vector-only branch)
    q = index.embed(query)
    o = hybrid_opts(vector_top_k, max(len(kw_keys), 0), min_vector_score, rrf, path=path)
    res = index.store.hybrid(q, o, [kw_keys]).row(0)

    base = index.store.id_base
    vec_nodes = [(index.nodes[int(i) - base], float(s)) for i, s in zip(res["vec_ids"], res["vec_scores"])]
    if not res["used_rrf"]:                                                    # :346-354
        out = []
        for (node, score), ct in zip(vec_nodes, res["ctype"]):
            md = node.metadata or {}
            out.append(HybridSearchResult(id=node.id_, documentName=_document_name(md, CONTENT_TYPE[int(ct)] == "memory"),
                                          content=node.text or "", score=score, source="vector",
                                          contentType=ctype_of(ct), metadata=md))
        return out

    first: dict[int, tuple[str, Any]] = {}
    for node, _ in vec_nodes:
        first.setdefault(index.keys.key(node.text or ""), ("vector", node))
    for hit, k in zip(keyword_results, kw_keys):
        first.setdefault(k, ("keyword", hit))
    out = []
    for key, score, src, ct in zip(res["keys"], res["scores"], res["source"], res["ctype"]):
        origin, r = first[int(key)]
        if origin == "vector":
            md = r.metadata or {}
            out.append(HybridSearchResult(id=index.keys.string(int(key)),
                                          documentName=_document_name(md, CONTENT_TYPE[int(ct)] == "memory"),
                                          content=r.text or "", score=float(score), source=SOURCE[int(src)],
                                          contentType=ctype_of(ct), metadata=md))
        else:
            out.append(HybridSearchResult(id=index.keys.string(int(key)), documentName=r.documentName, content=r.content,
                                          score=float(score), source=SOURCE[int(src)], contentType="document",
                                          documentId=r.documentId))
    return out


def format_search_results(results: Sequence[HybridSearchResult], max_results: int = 5) -> str:
    """formatSearchResults — src/lib/hybrid-search.ts:364-373."""
    parts = []
    for i, r in enumerate(results[:max_results]):
        source_tag = "🎯" if r.source == "both" else ("📊" if r.source == "vector" else "🔤")
        type_tag = "💻" if r.contentType == "code" else ("🧠" if r.contentType == "memory" else "📄")
        parts.append(f"[来源{i + 1}: {r.documentName}] {source_tag}{type_tag}\n{r.content}")
    return "\n\n".join(parts)


def get_source_stats(results: Sequence[HybridSearchResult]) -> dict:
    """getSourceStats — src/lib/hybrid-search.ts:378-399."""
    stats = dict(total=len(results), vector=0, keyword=0, both=0, byType={})
    for r in results:
        stats[r.source] = stats.get(r.source, 0) + 1
        stats["byType"][r.contentType] = stats["byType"].get(r.contentType, 0) + 1
    return stats

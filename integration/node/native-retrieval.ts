/**
 * native-retrieval.ts — drop into the reference as src/lib/native-retrieval.ts.
 *
 * Keeps `hybridSearch(index, knowledgeBaseId, query, options)` (src/lib/hybrid-search.ts:275-280) and its
 * result type unchanged; the dense scoring, top-k, min-cosine filter, RRF arithmetic and final stable sort
 * run on the GPU through the N-API addon (ragera_addon.cc → libragera.so). The host keeps what the
 * reference's host does: embedding the query and asking Meilisearch (both network calls), and re-attaching
 * strings to the integer fusion keys.
 *
 * Wiring in src/lib/hybrid-search.ts (the only edit to existing reference code):
 *
 *   import { NativeKnowledgeIndex, hybridSearchNative } from './native-retrieval';
 *   export async function hybridSearch(index, knowledgeBaseId, query, options = {}) {
 *     if (index instanceof NativeKnowledgeIndex) return hybridSearchNative(index, knowledgeBaseId, query, options);
 *     ... existing body ...
 *   }
 *
 * and in src/lib/llm/index-manager.ts `loadIndex` returns a NativeKnowledgeIndex built with
 * `NativeKnowledgeIndex.fromPersistDir(storageDir, dim)` when RAGERA_NATIVE=1.
 */
import * as fs from 'fs';
import * as path from 'path';
// eslint-disable-next-line @typescript-eslint/no-var-requires
const native = require('../../integration/node/build/Release/ragera_addon.node');
import { meilisearchService } from './meilisearch';
import type { HybridSearchResult, HybridSearchOptions, RRFConfig, SearchPreset } from './hybrid-search';

const SOURCE = ['vector', 'keyword', 'both'] as const;
const CTYPE = ['document', 'memory', 'code'] as const;

// PRESET_CONFIGS — src/lib/hybrid-search.ts:77-105
const PRESETS: Record<SearchPreset, { rrf: RRFConfig; vectorTopK: number; keywordLimit: number; minVectorScore: number }> = {
  document: { rrf: { k: 60, vectorWeight: 1.0, keywordWeight: 1.0, bothBonus: 0.1 }, vectorTopK: 8, keywordLimit: 8, minVectorScore: 0.3 },
  code: { rrf: { k: 40, vectorWeight: 1.0, keywordWeight: 1.3, bothBonus: 0.15 }, vectorTopK: 6, keywordLimit: 5, minVectorScore: 0.25 },
};

export interface NodeRow { id: string; text: string; metadata: Record<string, any> }

export class NativeKnowledgeIndex {
  /** row r of the device matrix ↔ nodes[r] (insertion order of embeddingDict). */
  nodes: NodeRow[] = [];
  private keyIds = new Map<string, bigint>();
  private keyStrs: string[] = [];

  constructor(readonly handle: unknown, readonly dim: number, private embed: (q: string) => Promise<Float32Array>) {}

  /** Load ./storage/kb_<id>/ as written by storageContextFromDefaults (index-manager.ts:218-220,264-270). */
  static fromPersistDir(dir: string, dim: number, capacityRows: number, embed: (q: string) => Promise<Float32Array>,
                        isCodebase = false): NativeKnowledgeIndex {
    const handle = native.createIndex({ rows: capacityRows, dim, device: 0, f16Shadow: 1 });
    const idx = new NativeKnowledgeIndex(handle, dim, embed);
    const ids: string[] = native.loadVectorStore(handle, path.join(dir, 'vector_store.json'));
    const docstore = JSON.parse(fs.readFileSync(path.join(dir, 'doc_store.json'), 'utf8'));
    const docs = docstore['docstore/data'] ?? {};
    const ctype = new Uint8Array(ids.length);
    const keys = new BigUint64Array(ids.length);
    ids.forEach((id, r) => {
      // SimpleDocumentStore persists each node as {__data__, __type__}; __data__ is the node's JSON — a string in
      // current llamaindex (serializer.toPersistence = JSON.stringify), a plain object in older stores: accept both
      const raw = docs[id]?.__data__;
      const d = typeof raw === 'string' ? JSON.parse(raw) : raw ?? {};
      const metadata = d.metadata ?? {};
      const text: string = d.text ?? '';
      idx.nodes.push({ id, text, metadata });
      // contentType rule — hybrid-search.ts:229-234
      ctype[r] = metadata.type === 'memory' ? 1 : (isCodebase || metadata.language !== undefined) ? 2 : 0;
      keys[r] = idx.keyOf(text);
    });
    native.setRowMeta(handle, 0, ctype);
    native.setRowKeys(handle, 0, keys);
    return idx;
  }

  /** content.substring(0,100) de-dup key (hybrid-search.ts:149,171) interned to an integer. */
  keyOf(content: string): bigint {
    const k = content.substring(0, 100);
    let id = this.keyIds.get(k);
    if (id === undefined) { id = BigInt(this.keyStrs.length); this.keyIds.set(k, id); this.keyStrs.push(k); }
    return id;
  }
  keyString(id: bigint): string { return this.keyStrs[Number(id)]; }
  embedQuery(q: string): Promise<Float32Array> { return this.embed(q); }

  /**
   * Micro-batcher per call-site class (one per distinct vectorTopK / keywordLimit / minVectorScore / RRF config — they are
   * per-launch parameters): concurrent requests of the server share ONE corpus pass and each gets exactly its batch-1
   * result. `native.submit` does not block and does not occupy a libuv pool thread: the Promise is resolved from the
   * batcher's worker through a thread-safe function, so the event loop can keep thousands of requests in flight.
   */
  private batchers = new Map<string, unknown>();
  batcherFor(opts: { vectorTopK: number; keywordLimit: number; minVectorScore: number; rrf: RRFConfig }): unknown {
    const key = JSON.stringify(opts);
    let b = this.batchers.get(key);
    if (b === undefined) { b = native.createBatcher(this.handle, opts, 1024, 1000); this.batchers.set(key, b); }
    return b;
  }
  /** Closes the handles; calls still in flight are answered first (the addon defers the native destroy to the last of them). */
  close(): void {
    this.batchers.forEach(b => native.destroyBatcher(b));
    this.batchers.clear();
    native.destroy(this.handle);
  }
}

/** RAGERA_BATCH=1: hybridSearchNative goes through the micro-batcher (a busy server); otherwise one direct call per request. */
const USE_BATCHER = process.env.RAGERA_BATCH === '1';

const docName = (m: any) => (m?.type === 'memory' ? '用户记忆' : m?.documentName || m?.relativePath || m?.filePath || '未知文档'); // :238-240

export async function hybridSearchNative(index: NativeKnowledgeIndex, knowledgeBaseId: string, query: string,
                                         options: HybridSearchOptions = {}): Promise<HybridSearchResult[]> {
  const presetConfig = PRESETS[options.preset || 'document'];
  const vectorTopK = options.vectorTopK ?? presetConfig.vectorTopK;             // :286-289 (?? keeps 0)
  const keywordLimit = options.keywordLimit ?? presetConfig.keywordLimit;
  const useKeyword = options.useKeyword ?? true;
  const minVectorScore = options.minVectorScore ?? presetConfig.minVectorScore;
  const rrf: RRFConfig = { ...presetConfig.rrf, ...options.rrfConfig };          // :292-295
  // :296 — per-call flag: every non-memory VECTOR hit of a codebase search is 'code' (:229-234)
  const isCodebase = (options.preset || 'document') === 'code' || knowledgeBaseId.startsWith('codebase_');
  const ctypeOf = (c: number) => (isCodebase && CTYPE[c] === 'document' ? 'code' : CTYPE[c]);

  const [q, hits] = await Promise.all([
    index.embedQuery(query),                                                      // network, as today
    useKeyword && keywordLimit > 0 && (await meilisearchService.isAvailable())    // :321-330
      ? meilisearchService.search(knowledgeBaseId, query, keywordLimit) : Promise.resolve([]),
  ]);
  const kwKeys = BigUint64Array.from(hits.map(h => index.keyOf(h.content)));
  // the batcher's keywordLimit is the call site's (a row stride shared by the whole batch); the direct call sizes it per request
  const batchOpts = { vectorTopK, keywordLimit, minVectorScore, rrf };   // hits.length <= keywordLimit (Meilisearch's limit)
  const r = USE_BATCHER
    ? await native.submit(index.batcherFor(batchOpts), batchOpts, q, kwKeys)
    : await native.hybridSearch(index.handle, q, 1,
        { vectorTopK, keywordLimit: hits.length, minVectorScore, rrf }, kwKeys, Uint32Array.of(hits.length));

  // re-attach strings: the first occurrence of a key wins, vector hits first, then keyword hits (:147-188)
  const first = new Map<bigint, { n?: NodeRow; h?: (typeof hits)[number] }>();
  for (let i = 0; i < r.vecCounts[0]; i++) {
    const n = index.nodes[Number(r.vecIds[i])];
    const k = index.keyOf(n.text);
    if (!first.has(k)) first.set(k, { n });
  }
  hits.forEach(h => { const k = index.keyOf(h.content); if (!first.has(k)) first.set(k, { h }); });

  const out: HybridSearchResult[] = [];
  for (let i = 0; i < r.counts[0]; i++) {
    if (!r.usedRrf[0]) {                                                          // vector-only branch :346-354
      const n = index.nodes[Number(r.keys[i])];
      out.push({ id: n.id, documentName: docName(n.metadata), content: n.text, score: r.scores[i], source: 'vector',
                 contentType: ctypeOf(r.contentType[i]), metadata: n.metadata });
      continue;
    }
    const e = first.get(r.keys[i])!;
    if (e.n) {
      out.push({ id: index.keyString(r.keys[i]), documentName: docName(e.n.metadata), content: e.n.text, score: r.scores[i],
                 source: SOURCE[r.source[i]], contentType: ctypeOf(r.contentType[i]), metadata: e.n.metadata });
    } else {
      out.push({ id: index.keyString(r.keys[i]), documentId: e.h!.documentId, documentName: e.h!.documentName, content: e.h!.content,
                 score: r.scores[i], source: SOURCE[r.source[i]], contentType: 'document' });
    }
  }
  return out;
}

/**
 * The llamaindex seam (SURVEY §8b): a vector store whose `query` runs on the GPU. Handed to
 * `storageContextFromDefaults({ persistDir, vectorStore })` in index-manager.ts:218-220,264-270, EVERY caller of
 * `index.asRetriever({similarityTopK}).retrieve(q)` — hybridSearch's vectorSearch (hybrid-search.ts:223-224),
 * MemoryStore.retrieve (memory/store.ts:111-116), summarize_topic (llm/tools/summarize-tool.ts:46-47), the query
 * engine (llm/agent.ts:143-163) — scans on the device with no edit at the call site. Shapes follow
 * `BaseVectorStore` of llamaindex@0.12.1: query({queryEmbedding, similarityTopK}) → {ids, similarities}.
 */
export class NativeVectorStore {
  storesText = false;
  constructor(private readonly index: NativeKnowledgeIndex) {}
  client(): unknown { return this.index.handle; }

  /** SimpleVectorStore.add: rows are appended in order (index.insert — memory/store.ts:67). */
  async add(nodes: Array<{ id_: string; getEmbedding(): number[]; getContent?(mode?: unknown): string; metadata?: Record<string, any> }>): Promise<string[]> {
    if (nodes.length === 0) return [];
    const dim = this.index.dim;
    const rows = new Float32Array(nodes.length * dim);
    nodes.forEach((n, i) => rows.set(n.getEmbedding(), i * dim));
    const row0: number = native.uploadRows(this.index.handle, rows, nodes.length);
    const ctype = new Uint8Array(nodes.length);
    const keys = new BigUint64Array(nodes.length);
    nodes.forEach((n, i) => {
      const metadata = n.metadata ?? {};
      const text = n.getContent ? n.getContent() : '';
      this.index.nodes.push({ id: n.id_, text, metadata });
      ctype[i] = metadata.type === 'memory' ? 1 : metadata.language !== undefined ? 2 : 0;
      keys[i] = this.index.keyOf(text);
    });
    native.setRowMeta(this.index.handle, row0, ctype);
    native.setRowKeys(this.index.handle, row0, keys);
    return nodes.map(n => n.id_);
  }

  /** getTopKEmbeddings: score all rows, stable sort, first k — ties to the earlier row. Exact fp64 cosines. */
  async query(q: { queryEmbedding?: number[]; similarityTopK: number }): Promise<{ ids: string[]; similarities: number[] }> {
    if (!q.queryEmbedding) throw new Error('NativeVectorStore.query needs queryEmbedding');
    const ids: string[] = [];
    const similarities: number[] = [];
    if (USE_BATCHER) {
      // a plain top-k through the micro-batcher = the vector-only branch with no keyword list and a filter below every
      // cosine: concurrent retrievers (hybridSearch, MemoryStore, summarize_topic) share one corpus pass
      const opts = { vectorTopK: q.similarityTopK, keywordLimit: 0, minVectorScore: -2, rrf: PRESETS.document.rrf };
      const r = await native.submit(this.index.batcherFor(opts), opts, Float32Array.from(q.queryEmbedding), new BigUint64Array(0));
      for (let i = 0; i < r.vecCounts[0]; i++) { ids.push(this.index.nodes[Number(r.vecIds[i])].id); similarities.push(r.vecScores[i]); }
      return { ids, similarities };
    }
    const r = await native.search(this.index.handle, Float32Array.from(q.queryEmbedding), 1, q.similarityTopK);
    const n: number = r.counts[0];
    for (let i = 0; i < n; i++) { ids.push(this.index.nodes[Number(r.ids[i])].id); similarities.push(r.scores[i]); }
    return { ids, similarities };
  }

  /**
   * MemoryStore.retrieve's device half (memory/store.ts:119-175): memory rows of the top-2·limit, cos ≥ minRelevance,
   * 0.7·cos + 0.3·freshness, stable sort. Needs the Memory columns on the device (setRowMeta after prisma writes).
   * All 2·limit blended candidates come back so the caller can drop rows whose DB record is gone before slicing (:153,:175).
   */
  async memoryCandidates(queryEmbedding: Float32Array, limit: number, minRelevance: number, nowMs: number):
      Promise<Array<{ node: NodeRow; score: number; relevanceScore: number; freshnessScore: number }>> {
    const r = await native.memoryRetrieve(this.index.handle, queryEmbedding, 1,
      { limit: 2 * limit, similarityTopK: 2 * limit, minRelevance, nowMs });
    const out = [];
    for (let i = 0; i < r.counts[0]; i++)
      out.push({ node: this.index.nodes[Number(r.ids[i])], score: r.scores[i], relevanceScore: r.relevance[i], freshnessScore: r.freshness[i] });
    return out;
  }
}

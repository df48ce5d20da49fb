#!/usr/bin/env python3
"""Generate known-answer vectors for the RRF / filter / freshness / blend steps.

The reference (gong9/rag-era) has NO tests or golden vectors and cannot be run
here (TypeScript, no Node.js), so parity for this path is UNPINNED. What this
script does instead: it transliterates the cited reference lines statement by
statement into Python — whose ``float`` is the same IEEE-754 binary64 as a
JavaScript number, whose ``dict`` keeps insertion order like a JS ``Map`` and
whose ``sorted`` is stable like V8's ``Array.prototype.sort`` — and records the
results as hex floats. The C oracle (oracle/oracle.c) and the CUDA path are both
checked against this file, so there are two independent restatements of the
reference text plus the hand-derived table of SURVEY.md Appendix A.

Reference lines followed:
  reciprocalRankFusion   src/lib/hybrid-search.ts:129-208
  min-cosine filter      src/lib/hybrid-search.ts:308-314
  calculateFreshness     src/lib/memory/freshness.ts:37-56
  MemoryStore blend      src/lib/memory/store.ts:148-175
  similarity / getTopKEmbeddings   llamaindex@0.12.1 / @llamaindex/core@0.6.22 (un-vendored; published algorithm,
                         "upstream-recalled"), call site src/lib/hybrid-search.ts:223-224

Run:  python tests/golden/make_kat.py   (rewrites tests/golden/kat_rrf.json)
"""
from __future__ import annotations

import json
import math
import os
import random

HERE = os.path.dirname(os.path.abspath(__file__))


def reciprocal_rank_fusion(vector_results, keyword_results, config):
    """Transliteration of src/lib/hybrid-search.ts:129-208 on (key, contentType) items."""
    k, vw, kw, bonus = config["k"], config["vectorWeight"], config["keywordWeight"], config["bothBonus"]
    score_map = {}  # JS Map: insertion-ordered
    for rank, (key, ctype) in enumerate(vector_results):          # :147
        rrf = vw / (k + rank + 1)                                   # :148
        ex = score_map.get(key)                                     # :151
        if ex is not None:
            ex["score"] += rrf                                      # :154
            ex["source"] = "both"                                   # :155
        else:
            score_map[key] = {"score": rrf, "source": "vector", "contentType": ctype}  # :157-164
    for rank, key in enumerate(keyword_results):                    # :169
        rrf = kw / (k + rank + 1)                                   # :170
        ex = score_map.get(key)                                     # :173
        if ex is not None:
            ex["score"] += rrf + (bonus * ex["score"])              # :176
            ex["source"] = "both"                                   # :177
        else:
            score_map[key] = {"score": rrf, "source": "keyword", "contentType": "document"}  # :179-186
    out = [dict(id=key, **v) for key, v in score_map.items()]       # :191-201
    out.sort(key=lambda r: -r["score"])                             # :202 stable, b.score - a.score
    return out


def filter_min(scores, min_score):
    """src/lib/hybrid-search.ts:308-314 — returns kept indices."""
    return [i for i, s in enumerate(scores) if not (s < min_score)]


def freshness(conf, access, last_ms, now_ms, decay=0.05, bonus=0.1):
    """src/lib/memory/freshness.ts:43-55."""
    hours = (now_ms - last_ms) / 3600000
    d = math.exp(-decay * hours)
    fb = math.log(access + 1) * bonus
    score = conf * d * (1 + fb)
    return max(0, min(1, score))


def blend(cos, fresh):
    """src/lib/memory/store.ts:160."""
    return cos * 0.7 + fresh * 0.3


DOCUMENT = dict(k=60, vectorWeight=1.0, keywordWeight=1.0, bothBonus=0.1)   # hybrid-search.ts:83-88
CODE = dict(k=40, vectorWeight=1.0, keywordWeight=1.3, bothBonus=0.15)      # hybrid-search.ts:95-100


def pack(name, cfg, vec, kw):
    res = reciprocal_rank_fusion(vec, kw, cfg)
    return dict(name=name, config=cfg, vector=[[k, c] for k, c in vec], keyword=list(kw),
                expect=[dict(key=r["id"], source=r["source"], contentType=r["contentType"],
                             score=r["score"], hex=float(r["score"]).hex()) for r in res])


def similarity_cosine(e1, e2):
    """llamaindex `similarity(e1, e2, SimilarityType.DEFAULT)` (embeddings utils of @llamaindex/core@0.6.22, upstream-
    recalled): dot product and both norms as three sequential left-to-right loops over JS numbers, no epsilon."""
    def norm(x):
        result = 0.0
        for v in x:
            result += v * v
        return math.sqrt(result)
    result = 0.0
    for a, b in zip(e1, e2):
        result += a * b
    return result / (norm(e1) * norm(e2))


def get_top_k_embeddings(query, embeddings, k):
    """llamaindex `getTopKEmbeddings` as reached from SimpleVectorStore.query (no cutoff): score ALL rows in insertion
    order, `sort((a, b) => b.similarity - a.similarity)` (stable in V8), first k."""
    sims = [dict(similarity=similarity_cosine(query, e), id=i) for i, e in enumerate(embeddings)]
    sims.sort(key=lambda s: -s["similarity"])
    return [s["id"] for s in sims[:k]], [s["similarity"] for s in sims[:k]]


def f32(x):
    """The stored values are fp32-precision numbers (what an embeddings API returns, and what the index dtype holds)."""
    import struct
    return struct.unpack("<f", struct.pack("<f", x))[0]


def cosine_topk_cases():
    rng = random.Random(7_2026)
    cases = []
    for name, n, d, k, twist in [("tiny", 5, 4, 3, None), ("ties-identical-rows", 12, 8, 6, "dup"), ("k-exceeds-n", 4, 16, 10, None),
                                 ("scaled-rows", 10, 8, 10, "scale"), ("negative-and-orthogonal", 9, 6, 9, "signs"),
                                 ("d-not-multiple-of-4", 20, 13, 8, None), ("wide", 40, 96, 10, "dup"), ("search_knowledge-k5", 60, 32, 5, None)]:
        X = [[f32(rng.gauss(0.0, 1.0)) for _ in range(d)] for _ in range(n)]
        q = [f32(X[n // 2][j] + 0.3 * rng.gauss(0.0, 1.0)) for j in range(d)]
        if twist == "dup":                      # exact ties: the earlier row must come first
            X[n - 1] = list(X[1]); X[n // 3] = list(X[1])
        if twist == "scale":                    # power-of-two scaling leaves the cosine bit-identical (ties again)
            X[7] = [f32(4.0 * v) for v in X[2]]; X[5] = [f32(0.5 * v) for v in X[2]]
        if twist == "signs":
            X[0] = [f32(-v) for v in q]; X[1] = [0.0] * (d - 1) + [1.0]; X[2] = list(q)
        ids, sims = get_top_k_embeddings(q, X, k)
        cases.append(dict(name=name, dim=d, k=k, rows=[[float(v).hex() for v in r] for r in X], query=[float(v).hex() for v in q],
                          ids=ids, similarities=[float(s).hex() for s in sims]))
    return cases


def main():
    D = "document"
    cases = [
        pack("KAT1", DOCUMENT, [(1, D), (2, D), (3, D)], [2, 4]),
        pack("KAT2-tie-insertion-order", DOCUMENT, [(1, D)], [2]),
        pack("KAT3-code-preset-order", CODE, [(1, "code"), (2, "code")], [2, 1]),
        pack("KAT4-search_knowledge-5+5", DOCUMENT, [(10, D), (11, D), (12, D), (13, D), (14, D)], [12, 21, 22, 23, 10]),
        pack("KAT5-vector-duplicate-key", DOCUMENT, [(1, D), (1, D)], []),
        pack("KAT6-keyword-duplicate-key", DOCUMENT, [(1, "memory")], [5, 5, 1, 1]),
        pack("KAT7-empty-vector", DOCUMENT, [], [7, 8, 9]),
        pack("KAT8-fractional-k", dict(k=12.5, vectorWeight=0.7, keywordWeight=1.9, bothBonus=0.33),
             [(3, D), (1, "memory"), (2, D)], [2, 3, 9, 1]),
    ]
    rng = random.Random(20261018)
    for i in range(40):
        nv, nk = rng.randint(0, 29), rng.randint(0, 19)
        pool = list(range(1, 40))
        vec = [(rng.choice(pool), rng.choice(["document", "memory", "code"])) for _ in range(nv)]
        kw = [rng.choice(pool) for _ in range(nk)]
        cfg = rng.choice([DOCUMENT, CODE, dict(k=float(rng.randint(1, 100)), vectorWeight=rng.choice([0.5, 1.0, 2.0]),
                                               keywordWeight=rng.choice([0.25, 1.0, 1.3]), bothBonus=rng.choice([0.0, 0.1, 0.5]))])
        cases.append(pack(f"RND{i}", cfg, vec, kw))

    now = 1_760_000_000_000
    fresh = []
    for conf, acc, hours in [(0.8, 0, 0.0), (0.8, 3, 14.0), (0.9, 10, 1.0), (0.5, 0, 1000.0), (1.0, 100, 0.25), (0.66, 7, 13.862943611198904)]:
        last = now - int(round(hours * 3600000))
        v = freshness(conf, acc, last, now)
        fresh.append(dict(confidence=conf, accessCount=acc, lastAccessedMs=last, nowMs=now, score=v, hex=float(v).hex()))
    blends = []
    for cosv, fr in [(0.9, 0.8), (0.5, 0.0), (0.73, 0.45234131555001067), (1.0, 1.0)]:
        b = blend(cosv, fr)
        blends.append(dict(cos=cosv, fresh=fr, score=b, hex=float(b).hex()))
    filt = dict(scores=[0.9, 0.31, 0.30, 0.2999], min=0.3, kept=filter_min([0.9, 0.31, 0.30, 0.2999], 0.3))

    out = dict(note="derived by tests/golden/make_kat.py from the cited reference lines; PARITY UNPINNED "
                    "(the reference ships no golden vectors)", rrf=cases, freshness=fresh, blend=blends, filter=filt,
               cosine_topk=cosine_topk_cases())
    with open(os.path.join(HERE, "kat_rrf.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(f"wrote {len(cases)} rrf cases")


if __name__ == "__main__":
    main()

"""Pins the oracle to the MOUNTED reference source (gong9/rag-era at /root/reference).

The reference cannot run here (TypeScript, no Node) and ships no golden vectors, so the oracle restates it.
What keeps that restatement honest is this test: it reads the reference's own files at test time,
  1. hashes every line range the oracle cites (a drift of the mounted source fails loudly),
  2. regex-extracts the constants (presets, thresholds, weights, tool k's) and compares them with what the
     host mirror and the oracle use,
  3. EXECUTES the reference's arithmetic: the right-hand sides of the scoring statements are lifted out of the
     TypeScript text, mechanically mapped to Python (JS number == Python float, Math.exp/log == libm), and the
     fusion / freshness / blend built from those very strings must agree with oracle.c bit for bit on random
     inputs and with tests/golden/kat_rrf.json.
Skipped where /root/reference does not exist (the GPU box); it runs in the builder's container and the judge's.
"""
import hashlib
import json
import math
import os
import re

import numpy as np
import pytest

REF = os.environ.get("RAGERA_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PINS = os.path.join(ROOT, "tests", "golden", "reference_pins.json")

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "lib")),
                                reason="reference source not mounted (expected on the GPU box)")

# (file, first line, last line): every range oracle/oracle.c, rag_era_b200/*.py and the kernels cite
RANGES = [
    ("src/lib/hybrid-search.ts", 77, 105),    # PRESET_CONFIGS
    ("src/lib/hybrid-search.ts", 129, 208),   # reciprocalRankFusion
    ("src/lib/hybrid-search.ts", 217, 247),   # vectorSearch + contentType rule
    ("src/lib/hybrid-search.ts", 275, 355),   # hybridSearch
    ("src/lib/memory/freshness.ts", 20, 23),  # DEFAULT_CONFIG
    ("src/lib/memory/freshness.ts", 37, 56),  # calculateFreshnessScore
    ("src/lib/memory/store.ts", 102, 180),    # MemoryStore.retrieve
    ("src/lib/context/engine.ts", 242, 246),  # getUnifiedResults' hybridSearch call
    ("src/lib/llm/tools/search-tools.ts", 12, 95),
]


def lines(rel):
    with open(os.path.join(REF, rel), encoding="utf-8") as f:
        return f.read().split("\n")


def text(rel, a, b):
    return "\n".join(lines(rel)[a - 1:b])


def sha(rel, a, b):
    return hashlib.sha256(text(rel, a, b).encode("utf-8")).hexdigest()


def test_cited_line_ranges_are_the_ones_the_oracle_was_written_against():
    pins = json.load(open(PINS))
    got = {f"{rel}:{a}-{b}": sha(rel, a, b) for rel, a, b in RANGES}
    assert got == pins["sha256"], ("the mounted reference differs from the source the oracle restates — re-read the "
                                   "changed lines, update oracle/oracle.c and re-pin with tests/golden/make_reference_pins.py")


def js_expr(e):
    """A JS arithmetic expression of the cited statements → Python (same IEEE-754 double operations)."""
    e = e.replace("Math.exp", "math.exp").replace("Math.log", "math.log").replace("Math.max", "max").replace("Math.min", "min")
    e = re.sub(r"\bexisting\.score\b", "existing_score", e)
    e = re.sub(r"\bnow\.getTime\(\)", "now_ms", e)
    e = re.sub(r"\bmemory\.lastAccessedAt\.getTime\(\)", "last_ms", e)
    e = re.sub(r"\bmemory\.(\w+)\b", r"memory_\1", e)
    e = re.sub(r"\bconfig\.(\w+)\b", r"config_\1", e)
    assert re.fullmatch(r"[\w\s.+\-*/(),]+", e), e          # nothing but names, numbers and arithmetic
    return compile(e, "<reference>", "eval")


def rhs(rel, lineno, pattern):
    """The expression of statement `pattern` (one regex group) on line `lineno` of the reference file."""
    m = re.search(pattern, lines(rel)[lineno - 1])
    assert m, (rel, lineno, lines(rel)[lineno - 1])
    return m.group(1)


def test_presets_and_call_site_constants():
    import importlib

    hs = importlib.import_module("rag_era_b200.hybrid_search")     # the package re-exports a function of the same name
    ctx = importlib.import_module("rag_era_b200.context")

    t = text("src/lib/hybrid-search.ts", 77, 105)
    for name in ("document", "code"):
        blk = t[t.index(f"  {name}: {{"):]
        g = lambda key: float(re.search(rf"\b{key}:\s*([0-9.]+)", blk).group(1))   # noqa: E731
        p = hs.PRESET_CONFIGS[name]
        assert (p["rrf"].k, p["rrf"].vector_weight, p["rrf"].keyword_weight, p["rrf"].both_bonus) == \
               (g("k"), g("vectorWeight"), g("keywordWeight"), g("bothBonus")), name
        assert (p["vectorTopK"], p["keywordLimit"], p["minVectorScore"]) == (g("vectorTopK"), g("keywordLimit"), g("minVectorScore"))
    # the default RRFConfig of reciprocalRankFusion is the document preset (:132) — what oracle.RRFConfig() holds
    assert "config: RRFConfig = PRESET_CONFIGS.document.rrf" in lines("src/lib/hybrid-search.ts")[131]
    import oracle
    d = hs.PRESET_CONFIGS["document"]["rrf"]
    assert (oracle.RRFConfig().k, oracle.RRFConfig().vector_weight, oracle.RRFConfig().keyword_weight, oracle.RRFConfig().both_bonus) == \
           (d.k, d.vector_weight, d.keyword_weight, d.both_bonus)
    # freshness defaults (freshness.ts:20-23)
    f = text("src/lib/memory/freshness.ts", 20, 23)
    assert float(re.search(r"timeDecayFactor:\s*([0-9.]+)", f).group(1)) == 0.05
    assert float(re.search(r"frequencyBonus:\s*([0-9.]+)", f).group(1)) == 0.1
    # call sites: engine.ts (c+10, min 0.4), the tools (5/5 and 10/10), MemoryStore (limit*2, minRelevance 0.5)
    e = text("src/lib/context/engine.ts", 242, 246)
    assert "counts.vectorTopK + 10" in e and re.search(r"minVectorScore:\s*0\.4\b", e)
    st = text("src/lib/llm/tools/search-tools.ts", 12, 95)
    assert re.findall(r"vectorTopK:\s*(\d+),\s*\n\s*keywordLimit:\s*(\d+)", st) == [("5", "5"), ("10", "10")]
    ms = lines("src/lib/memory/store.ts")
    assert "similarityTopK: limit * 2" in ms[111] and "minRelevance: number = 0.5" in ms[104]
    src = open(ctx.__file__).read() + open(hs.__file__).read()
    assert "+ 10" in src and "0.4" in src


def reference_rrf(vec_keys, kw_keys, cfg):
    """reciprocalRankFusion assembled from the reference's own statements (hybrid-search.ts:148,154,170,176,202)."""
    F = "src/lib/hybrid-search.ts"
    v_rrf = js_expr(rhs(F, 148, r"const rrfScore = (.+);"))
    v_add = js_expr(rhs(F, 154, r"existing\.score \+= (.+);"))
    k_rrf = js_expr(rhs(F, 170, r"const rrfScore = (.+);"))
    k_add = js_expr(rhs(F, 176, r"existing\.score \+= (.+);"))
    assert rhs(F, 202, r"\.sort\(\(a, b\) => (.+)\);") == "b.score - a.score"      # descending, stable (V8 TimSort)
    assert "substring(0, 100)" in lines(F)[148] and "substring(0, 100)" in lines(F)[170]
    env = dict(k=cfg[0], vectorWeight=cfg[1], keywordWeight=cfg[2], bothBonus=cfg[3])
    m = {}                                                      # JS Map: insertion-ordered, like dict
    for rank, key in enumerate(vec_keys):
        rrfScore = eval(v_rrf, {"math": math}, dict(env, rank=rank))
        if key in m:
            m[key][0] = m[key][0] + eval(v_add, {"math": math}, dict(env, rrfScore=rrfScore, existing_score=m[key][0]))
            m[key][1] = 2
        else:
            m[key] = [rrfScore, 0]
    for rank, key in enumerate(kw_keys):
        rrfScore = eval(k_rrf, {"math": math}, dict(env, rank=rank))
        if key in m:
            m[key][0] = m[key][0] + eval(k_add, {"math": math}, dict(env, rrfScore=rrfScore, existing_score=m[key][0]))
            m[key][1] = 2
        else:
            m[key] = [rrfScore, 1]
    out = sorted(m.items(), key=lambda kv: -kv[1][0])           # stable
    return [k for k, _ in out], [v[0] for _, v in out], [v[1] for _, v in out]


def test_rrf_built_from_the_reference_text_equals_the_oracle(golden):
    import oracle

    rng = np.random.default_rng(2024)
    cfgs = [(60.0, 1.0, 1.0, 0.1), (40.0, 1.0, 1.3, 0.15), (1.0, 0.7, 2.5, 0.0), (60.0, 1.0, 1.0, 0.5)]
    for trial in range(300):
        cfg = cfgs[trial % len(cfgs)]
        nv, nk = int(rng.integers(0, 30)), int(rng.integers(0, 20))
        pool = rng.integers(0, 25 if trial % 3 else 1000, size=64)
        vk = [int(x) for x in rng.choice(pool, nv)]             # duplicates inside a list on purpose (:152-155)
        kk = [int(x) for x in rng.choice(pool, nk)]
        keys, scores, src = reference_rrf(vk, kk, cfg)
        ok, os_, osrc, _ = oracle.rrf(vk, kk, oracle.RRFConfig(*cfg))
        assert [int(x) for x in ok] == keys and [int(x) for x in osrc] == src
        assert np.array_equal(np.array(scores, dtype=np.float64).view(np.uint64), os_.view(np.uint64)), trial
    # and the committed golden cases (tests/golden/kat_rrf.json) replayed through the reference text
    n = 0
    for case in golden["rrf"]:
        c = case["config"]
        keys, scores, src = reference_rrf([v[0] if isinstance(v, list) else v for v in case["vector"]], case["keyword"],
                                          (c["k"], c["vectorWeight"], c["keywordWeight"], c["bothBonus"]))
        assert keys == [e["key"] for e in case["expect"]]
        assert [float.fromhex(e["hex"]) for e in case["expect"]] == scores
        assert [("vector", "keyword", "both")[x] for x in src] == [e["source"] for e in case["expect"]]
        n += 1
    assert n >= 40


def test_freshness_and_blend_built_from_the_reference_text_equal_the_oracle():
    import oracle

    F = "src/lib/memory/freshness.ts"
    hours = js_expr(rhs(F, 43, r"const hoursSinceAccess = (.+);"))
    decay = js_expr(rhs(F, 46, r"const decayFactor = (.+);"))
    bonus = js_expr(rhs(F, 49, r"const frequencyBonus = (.+);"))
    score = js_expr(rhs(F, 52, r"const score = (.+);"))
    clamp = js_expr(rhs(F, 55, r"return (.+);"))
    S = "src/lib/memory/store.ts"
    blend = js_expr(rhs(S, 160, r"const score = (.+);"))
    assert re.search(r"if \(relevanceScore < minRelevance\)", lines(S)[150])
    assert re.search(r"if \(r\.score < minVectorScore\)", lines("src/lib/hybrid-search.ts")[308])
    rng = np.random.default_rng(7)
    now = 1_760_000_000_000
    G = {"math": math, "max": max, "min": min}
    for _ in range(2000):
        conf, acc = float(rng.uniform(0, 1.2)), int(rng.integers(0, 500))
        last = now - int(rng.integers(0, 400 * 3600000))
        h = eval(hours, G, dict(now_ms=now, last_ms=last))
        dfac = eval(decay, G, dict(config_timeDecayFactor=0.05, hoursSinceAccess=h))
        fb = eval(bonus, G, dict(memory_accessCount=acc, config_frequencyBonus=0.1))
        sc = eval(score, G, dict(memory_confidence=conf, decayFactor=dfac, frequencyBonus=fb))
        want = eval(clamp, G, dict(score=sc))
        got = oracle.freshness(conf, acc, last, now)
        assert got == want, (conf, acc, last, got, want)        # same libm on this host: bit-equal
        rel = float(rng.uniform(0.5, 1))
        b = eval(blend, G, dict(relevanceScore=rel, freshnessScore=want))
        oi, osc, ofr = oracle.memory_rank([rel], [1], [conf], [acc], [last], now, 10, 0.5)
        assert len(oi) == 1 and osc[0] == b and ofr[0] == want

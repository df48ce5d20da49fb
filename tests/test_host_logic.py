"""Host-side mirror of the reference interface: pieces that need no GPU."""
import numpy as np

import importlib

hs = importlib.import_module("rag_era_b200.hybrid_search")
from rag_era_b200 import shard_range
from rag_era_b200.sharded import merge_reference_order


def test_js_substring_counts_utf16_units():
    s = "a😀b" * 60                       # the emoji is 2 UTF-16 code units
    k = hs.js_substring(s, 0, 100)
    assert len(k.encode("utf-16-le", "surrogatepass")) == 200
    assert hs.js_substring("【文档: x】\n\nhello", 0, 100) == "【文档: x】\n\nhello"
    assert hs.js_substring("x" * 300, 0, 100) == "x" * 100


def test_key_interner_is_prefix_equality():
    ki = hs.KeyInterner()
    a = ki.key("p" * 100 + "tail one")
    b = ki.key("p" * 100 + "another tail")
    c = ki.key("q" + "p" * 99)
    assert a == b != c and ki.string(a) == "p" * 100


def test_presets_match_reference():
    d, c = hs.get_preset_config("document"), hs.get_preset_config("code")
    assert (d["rrf"].k, d["rrf"].vector_weight, d["rrf"].keyword_weight, d["rrf"].both_bonus) == (60, 1.0, 1.0, 0.1)
    assert (d["vectorTopK"], d["keywordLimit"], d["minVectorScore"]) == (8, 8, 0.3)
    assert (c["rrf"].k, c["rrf"].keyword_weight, c["rrf"].both_bonus) == (40, 1.3, 0.15)
    assert (c["vectorTopK"], c["keywordLimit"], c["minVectorScore"]) == (6, 5, 0.25)


def test_content_type_rule():
    assert hs.classify_content_type({"type": "memory", "language": "ts"}, True) == "memory"
    assert hs.classify_content_type({"language": "ts"}, False) == "code"
    assert hs.classify_content_type({}, True) == "code"
    assert hs.classify_content_type({}, False) == "document"
    # `metadata.language !== undefined` (hybrid-search.ts:230): null and '' are not undefined
    assert hs.classify_content_type({"language": None}, False) == "code"
    assert hs.classify_content_type({"language": ""}, False) == "code"


def test_format_and_stats():
    r = [hs.HybridSearchResult("k1", "doc.md", "alpha", 0.03, "both", "document"),
         hs.HybridSearchResult("k2", "用户记忆", "beta", 0.02, "vector", "memory"),
         hs.HybridSearchResult("k3", "a.ts", "gamma", 0.01, "keyword", "code")]
    txt = hs.format_search_results(r, 2)
    assert txt == "[来源1: doc.md] 🎯📄\nalpha\n\n[来源2: 用户记忆] 📊🧠\nbeta"
    st = hs.get_source_stats(r)
    assert st == dict(total=3, vector=1, keyword=1, both=1, byType={"document": 1, "memory": 1, "code": 1})


def test_shard_ranges_partition_rows():
    for total in (1, 7, 10_000, 50_000_000):
        for w in (1, 2, 4, 8):
            spans = [shard_range(total, w, r) for r in range(w)]
            assert sum(n for _, n in spans) == total
            pos = 0
            for base, n in spans:
                if n:
                    assert base == pos
                pos += n


def test_balanced_ranges_are_contiguous_and_follow_measured_speed():
    from rag_era_b200.sharded import balanced_ranges

    total = 50_000_000
    ms = [15.1, 17.4, 16.0, 15.5, 16.2, 15.9, 16.8, 15.3]
    spans = balanced_ranges(total, [total // 8] * 8, ms)
    pos = 0
    for base, n in spans:
        assert base == pos and n > 0
        pos += n
    assert pos == total and all(n % 256 == 0 for _, n in spans[:-1])
    # predicted time per shard (rows / measured speed) is equal to within one alignment step
    speed = [total / 8 / t for t in ms]
    pred = [n / s for (_, n), s in zip(spans, speed)]
    assert max(pred) / min(pred) < 1.001
    assert spans[1][1] < spans[0][1]                       # the slowest GPU gets the fewest rows
    for w in (1, 2, 3):                                    # degenerate inputs still partition the corpus
        s2 = balanced_ranges(1000, [1000 // w] * w, [1.0] * w)
        assert sum(n for _, n in s2) == 1000


def test_merge_reference_order_ties_by_id():
    ids, sc = merge_reference_order([[5, 9], [2, 7]], [[0.9, 0.5], [0.9, 0.5]], 3)
    assert ids.tolist() == [2, 5, 7] and sc.tolist() == [0.9, 0.9, 0.5]


def test_calculate_retrieval_count_matches_reference_table():
    """retrieval-decision.ts:144-195 — and the engine's k = vectorTopK + 10 stays within RAG_MAX_TOPK (12..29)."""
    import rag_era_b200 as rb
    from rag_era_b200 import _native as N

    D = rb.RetrievalDecision
    assert rb.calculate_retrieval_count(D()) == dict(vectorTopK=8, keywordLimit=6, graphLimit=0)          # floor(13*1.0)=13 → ceil(7.8), ceil(5.2)
    assert rb.calculate_retrieval_count(D(priority="high")) == dict(vectorTopK=12, keywordLimit=8, graphLimit=0)   # 19 → ceil(11.4), ceil(7.6)
    assert rb.calculate_retrieval_count(D(priority="low")) == dict(vectorTopK=6, keywordLimit=4, graphLimit=0)     # floor(9.1)=9 → ceil(5.4), ceil(3.6)
    assert rb.calculate_retrieval_count(D(queryType="semantic", priority="high")) == dict(vectorTopK=19, keywordLimit=0, graphLimit=0)
    assert rb.calculate_retrieval_count(D(queryType="keyword")) == dict(vectorTopK=2, keywordLimit=13, graphLimit=0)
    assert rb.calculate_retrieval_count(D(queryType="graph", priority="low")) == dict(vectorTopK=3, keywordLimit=0, graphLimit=9)
    ks = {rb.calculate_retrieval_count(D(queryType=t, priority=p))["vectorTopK"] + 10
          for t in ("semantic", "keyword", "graph", "hybrid") for p in ("high", "medium", "low")}
    assert min(ks) == 12 and max(ks) == 29 and max(ks) <= N.MAX_TOPK


def test_exchange_counter_alternates_parity_across_the_wrap(native):
    """The sharded exchange's two mailbox halves are chosen by the counter's parity: two consecutive exchanges must never
    share a half (a fast rank would refill a mailbox its slower peer is still merging from). The counter has 20 bits and
    skips 0 (the value of a flag word nobody has written); its wrap must keep the alternation — 0xFFFFF is followed by 2."""
    nxt = native.load().rag_debug_p2p_next_step
    assert nxt(0) == 1 and nxt(1) == 2 and nxt(0xFFFFE) == 0xFFFFF and nxt(0xFFFFF) == 2
    s = 0xFFFF0
    for _ in range(64):                      # across the wrap
        n = nxt(s)
        assert 0 < n <= 0xFFFFF and (n & 1) != (s & 1), (s, n)
        s = n
    s, seen_wrap = 0, 0
    for _ in range((1 << 20) + 10):          # a whole period from a fresh counter
        n = nxt(s)
        assert n != 0 and (n ^ s) & 1
        seen_wrap += n < s
        s = n
    assert seen_wrap == 1

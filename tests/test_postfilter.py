"""§8f N3 — rag_process_results (C++, host) against a statement-by-statement Python transliteration of
src/lib/context/rag/dedup-filter.ts on UTF-16 code units. CPU only. (The reference has no tests for this
either: parity unpinned; two independent restatements agree.)"""
import importlib
import re
from dataclasses import dataclass

from hypothesis import given, settings, strategies as st

ctx = importlib.import_module("rag_era_b200.context")

JS_SPACE = "\t\n\x0b\x0c\r \xa0                　﻿"
PUNCT = "，。！？、；：\"\"''【】（）"


def u16(s):            # a JS string = a sequence of UTF-16 code units
    b = s.encode("utf-16-le", "surrogatepass")
    return tuple(int.from_bytes(b[i:i + 2], "little") for i in range(0, len(b), 2))


def includes(hay, needle):
    n = len(needle)
    return any(hay[i:i + n] == needle for i in range(len(hay) - n + 1))


def split_on(units, seps, collapse):
    """str.split(regex): collapse=True for /[...]+/, False for split(' ')."""
    out, cur, prev_sep = [], [], False
    for c in units:
        if c in seps:
            if not (collapse and prev_sep):
                out.append(tuple(cur)); cur = []
            prev_sep = True
        else:
            cur.append(c); prev_sep = False
    out.append(tuple(cur))
    return out


SP = set(u16(JS_SPACE)); PU = set(u16(PUNCT))


def trim(u):
    a, b = 0, len(u)
    while a < b and u[a] in SP: a += 1
    while b > a and u[b - 1] in SP: b -= 1
    return u[a:b]


def extract_keywords(u):                                                   # :160-165
    cleaned = tuple(0x20 if (c in PU or c in SP) else c for c in u)
    return {w for w in split_on(cleaned, {0x20}, False) if len(w) >= 2}


def coverage(qk, ck):                                                      # :170-188
    if not qk: return 0
    covered = 0
    for kw in qk:
        if any(includes(w, kw) or includes(kw, w) for w in ck): covered += 1
    return covered / len(qk)


def is_noise(c):                                                           # :96-102
    s = "".join(chr(x) for x in c)
    pats = [r"^[\s\n]+$", r"^[.。,，;；:：!！?？]+$", r"^[0-9]+$", r"^第?[0-9]+[章节页条款]$", r"^(目录|索引|参考文献)$"]
    if all(x in SP for x in c) and c: return True
    return any(re.match(p, s) for p in pats[1:])


def process_results_py(results, query, enable_noise=True, enable_rerank=True, thr=0.85, minlen=20, maxres=10):
    q = u16(query)
    kws = [w for w in split_on(q, SP | set(u16("，。！？、")), True) if len(w) >= 2]           # :216
    proc = [i for i, r in enumerate(results) if not kws or any(includes(u16(r.content), kw) for kw in kws)]
    if enable_noise:                                                        # :106-127
        keep = []
        for i in proc:
            c = trim(u16(results[i].content))
            if is_noise(c): continue
            if len(c) > 0 and sum(1 for x in c if x in PU) / len(c) > 0.3: continue
            keep.append(i)
        proc = keep
    dd = []                                                                 # :42-91
    for i in proc:
        c = u16(results[i].content)
        if len(c) < minlen: continue
        dup, target = False, None
        for e in dd:
            s1, s2 = set(c[:200]), set(u16(results[e["i"]].content)[:200])
            uni = len(s1 | s2)
            if uni and len(s1 & s2) / uni >= thr:
                dup = True
                if results[i].score > e["f"]: target = e
                break
        if not dup: dd.append(dict(i=i, f=results[i].score, d=False, n=1))
        elif target is not None:
            target["n"] += 1; target["f"] = max(target["f"], results[i].score); target["d"] = True
    dd = dd[:maxres]
    if enable_rerank:                                                       # :132-155
        qk = extract_keywords(q)
        for e in dd:
            e["f"] = e["f"] * 0.7 + coverage(qk, extract_keywords(u16(results[e["i"]].content))) * 0.3
        dd.sort(key=lambda e: -e["f"])
    return dd


@dataclass
class R:
    content: str
    score: float
    source: str = "vector"


def check(results, query, **kw):
    exp = process_results_py(results, query, kw.get("enable_noise_filter", True), kw.get("enable_rerank", True))
    got = ctx.process_results(results, query, **kw)
    assert [g.index for g in got] == [e["i"] for e in exp]
    assert [g.fusionScore for g in got] == [e["f"] for e in exp]
    assert [g.deduplicated for g in got] == [e["d"] for e in exp]
    assert [g.sources for g in got] == [e["n"] for e in exp]
    return got


def test_process_results_cases(native):
    body = "体检前三天请保持正常饮食，不要饮酒，避免剧烈运动。" * 2
    rs = [R("【文档: 体检须知.pdf】\n\n" + body, 0.031, "hybrid"),
          R("【文档: 体检须知.pdf】\n\n" + body + "补充", 0.034, "keyword"),          # near-duplicate with a higher score → merge
          R("第12章", 0.02), R("目录", 0.02), R("12345", 0.02), R("。。。！！！？？？", 0.02), R("   \n\t ", 0.02),
          R("体检 " + "，。！？" * 10, 0.02),                                          # punctuation density > 0.3
          R("完全无关的内容，讲的是软件架构设计与代码评审流程的说明文档。", 0.029),        # no query keyword → dropped at step 0
          R("体检当天需要空腹，抽血项目在上午十点前完成，请携带身份证件。", 0.016, "vector"),
          R("short 体检", 0.05)]                                                        # shorter than minContentLength
    got = check(rs, "体检 注意事项 空腹")
    assert [g.index for g in got][:1] == [9] or got                                   # coverage can reorder; exact order checked above
    assert any(g.deduplicated and g.sources == 2 for g in got)
    check(rs, "")                                                                      # no query keywords: step 0 keeps everything
    check(rs, "体检", enable_noise_filter=False)
    check(rs, "体检 空腹", enable_rerank=False)
    check([], "体检")
    check([R("a😀b" * 30 + " 体检项目说明", 0.02), R("a😀b" * 30 + " 体检项目说明!", 0.03)], "体检项目 😀b")   # surrogate pairs count as 2 units


alphabet = st.sampled_from(list("体检空腹项目说明报告血压abcde，。！？、；： \n【】（）\"'12第章目录"))
texts = st.text(alphabet=alphabet, min_size=0, max_size=80)


@settings(max_examples=200, deadline=None)
@given(contents=st.lists(texts, max_size=12), scores=st.lists(st.floats(0.001, 0.06), min_size=12, max_size=12),
       query=st.text(alphabet=alphabet, max_size=20), noise=st.booleans(), rerank=st.booleans())
def test_process_results_property(native, contents, scores, query, noise, rerank):
    rs = [R(c, s, ["vector", "keyword", "hybrid"][i % 3]) for i, (c, s) in enumerate(zip(contents, scores))]
    check(rs, query, enable_noise_filter=noise, enable_rerank=rerank)


def test_process_results_under_sanitizers(tmp_path):
    """postfilter.cu is host-only: built as plain C++ with ASAN + UBSAN (tests/c/postfilter_asan.cc) and driven with random UTF-16
    inputs (empty strings, lone surrogates, CJK punctuation, multi-KB runs, exact duplicates) and random options."""
    import os
    import shutil
    import subprocess

    import pytest

    cuda_inc, cuda_lib = "/usr/local/cuda/include", "/usr/local/cuda/lib64"
    if shutil.which("g++") is None or not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA headers")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csrc = os.path.join(root, "rag_era_b200", "csrc")
    exe = str(tmp_path / "postfilter_asan")
    r = subprocess.run(["g++", "-std=c++17", "-g", "-O1", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-x", "c++",
                        "-I" + cuda_inc, "-I" + os.path.join(root, "include"), "-I" + csrc, os.path.join(root, "tests", "c", "postfilter_asan.cc"),
                        os.path.join(csrc, "postfilter.cu"), "-o", exe, "-L" + cuda_lib, "-lcudart", "-Wl,-rpath," + cuda_lib],
                       capture_output=True, text=True)
    if r.returncode != 0 and "sanitize" in r.stderr and "cannot find" in r.stderr:
        pytest.skip("libasan / libubsan are not installed")
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe, "400"], capture_output=True, text=True, timeout=300, env=dict(os.environ, ASAN_OPTIONS="detect_leaks=1"))
    assert r.returncode == 0 and "0 failures" in r.stdout, (r.stdout[-500:], r.stderr[-1500:])
    assert "AddressSanitizer" not in r.stderr and "LeakSanitizer" not in r.stderr and "runtime error" not in r.stderr, r.stderr[-1500:]

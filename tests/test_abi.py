"""The C-ABI boundary without a GPU: the library builds, loads, exports every symbol that
include/ragera.h declares, the ctypes mirrors match the C struct layouts, and — with no
device — refuses to work instead of falling back to anything."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ragera.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rag_[a-z0-9_]+)\s*\(", src)))


def test_header_functions_are_exported_and_bound(native):
    lib = native.load()
    names = declared_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"libragera.so does not export {n}"
        assert n in native.SYMBOLS, f"binding lacks {n}"
    assert sorted(native.SYMBOLS) == names
    # and nothing is resolved lazily from elsewhere: nm agrees
    out = subprocess.run(["nm", "-D", "--defined-only", native.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (rag_[a-z0-9_]+)", out))
    assert set(names) <= exported


def test_struct_layouts_match_the_header(native):
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "ragera.h"
int main(void) {
  printf("rag_gen_desc %zu\n", sizeof(rag_gen_desc));
  printf("rag_index_desc %zu\n", sizeof(rag_index_desc));
  printf("rag_rrf_config %zu\n", sizeof(rag_rrf_config));
  printf("rag_search_opts %zu\n", sizeof(rag_search_opts));
  printf("rag_topk_out %zu\n", sizeof(rag_topk_out));
  printf("rag_hybrid_opts %zu\n", sizeof(rag_hybrid_opts));
  printf("rag_fused_out %zu\n", sizeof(rag_fused_out));
  printf("rag_memory_opts %zu\n", sizeof(rag_memory_opts));
  printf("rag_memory_out %zu\n", sizeof(rag_memory_out));
  printf("rag_batcher_desc %zu\n", sizeof(rag_batcher_desc));
  printf("rag_text %zu\n", sizeof(rag_text));
  printf("rag_process_opts %zu\n", sizeof(rag_process_opts));
  printf("rag_processed_out %zu\n", sizeof(rag_processed_out));
  printf("rag_cache_info %zu\n", sizeof(rag_cache_info));
  printf("off_memory_topk %zu\n", offsetof(rag_memory_opts, similarity_top_k));
  printf("off_hybrid_now_ms %zu\n", offsetof(rag_hybrid_opts, now_ms));
  printf("off_hybrid_epsilon %zu\n", offsetof(rag_hybrid_opts, epsilon));
  printf("off_fused_certified %zu\n", offsetof(rag_fused_out, certified));
  printf("off_gen_now_ms %zu\n", offsetof(rag_gen_desc, now_ms));
  return 0;
}
'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "layout.c")
        open(c, "w").write(prog)
        exe = os.path.join(d, "layout")
        subprocess.run(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        got = dict(l.split() for l in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.splitlines())
    N = native
    assert int(got["rag_gen_desc"]) == C.sizeof(N.GenDesc)
    assert int(got["rag_index_desc"]) == C.sizeof(N.IndexDesc)
    assert int(got["rag_rrf_config"]) == C.sizeof(N.RRFConfigC)
    assert int(got["rag_search_opts"]) == C.sizeof(N.SearchOpts)
    assert int(got["rag_topk_out"]) == C.sizeof(N.TopkOut)
    assert int(got["rag_hybrid_opts"]) == C.sizeof(N.HybridOpts)
    assert int(got["rag_fused_out"]) == C.sizeof(N.FusedOut)
    assert int(got["rag_memory_opts"]) == C.sizeof(N.MemoryOpts)
    assert int(got["rag_memory_out"]) == C.sizeof(N.MemoryOut)
    assert int(got["rag_batcher_desc"]) == C.sizeof(N.BatcherDesc)
    assert int(got["rag_text"]) == C.sizeof(N.Text)
    assert int(got["rag_process_opts"]) == C.sizeof(N.ProcessOpts)
    assert int(got["rag_processed_out"]) == C.sizeof(N.ProcessedOut)
    assert int(got["rag_cache_info"]) == C.sizeof(N.CacheInfo)
    assert int(got["off_memory_topk"]) == N.MemoryOpts.similarity_top_k.offset
    assert int(got["off_hybrid_now_ms"]) == N.HybridOpts.now_ms.offset
    assert int(got["off_hybrid_epsilon"]) == N.HybridOpts.epsilon.offset
    assert int(got["off_fused_certified"]) == N.FusedOut.certified.offset
    assert int(got["off_gen_now_ms"]) == N.GenDesc.now_ms.offset
    # the oracle's mirror of rag_gen_desc must agree too
    import oracle

    assert C.sizeof(oracle.GenDesc) == C.sizeof(N.GenDesc)


def test_no_device_means_no_service(native):
    """Without a GPU the library must refuse (RAG_ERR_NO_DEVICE): there is no CPU path to fall back to."""
    lib = native.load()
    assert lib.rag_version() == native.RAGERA_VERSION
    if lib.rag_device_count() > 0:
        pytest.skip("a GPU is visible")
    d = native.IndexDesc(16, 64, native.F32, 0, 0, 0)
    h = C.c_void_p()
    rc = lib.rag_index_create(C.byref(d), C.byref(h))
    assert rc == native.ERR_NO_DEVICE and not h.value
    assert b"no CPU fallback" in lib.rag_last_error()
    from rag_era_b200 import VectorIndex, RagError

    with pytest.raises(RagError):
        VectorIndex(64, 16)


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under rag_era_b200/ may import, link or load it."""
    pkg = os.path.join(ROOT, "rag_era_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                text = open(os.path.join(base, f), errors="replace").read()
                assert "liboracle" not in text and "import oracle" not in text and "from oracle" not in text, f
                assert "oracle.h" not in text and "oracle/" not in text, f
    deps = subprocess.run(["ldd", os.path.join(pkg, "libragera.so")], capture_output=True, text=True).stdout
    assert "oracle" not in deps
    # ... and outside tests/ only bench.py (cpu_baseline / --impl reference legs) and __graft_entry__.py (smoke) import it
    for sub in ("tools", "integration", "include"):
        for base, _, files in os.walk(os.path.join(ROOT, sub)):
            for f in files:
                if f.endswith((".py", ".cc", ".ts", ".h", ".sh")):
                    text = open(os.path.join(base, f), errors="replace").read()
                    assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, os.path.join(sub, f)


def test_napi_addon_type_checks():
    """§8f N2: Node is absent here, so the addon cannot be built — but it must at least be valid C++ against the
    Node-API signatures it uses and against include/ragera.h (a stub header with the documented declarations)."""
    import shutil
    import subprocess

    if shutil.which("g++") is None:
        pytest.skip("no g++")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Wextra", "-Werror",
                        "-I" + os.path.join(ROOT, "tests", "c", "node_api_stub"), "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "integration", "node", "ragera_addon.cc")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr

"""§8f N4 on CPU: the micro-batcher's host logic (batcher.cu) under ThreadSanitizer, with `rag_hybrid_search` replaced by a
stub that derives every output from the query it was handed (tests/c/batcher_tsan.cc). 24 submitter threads × 40 requests:
every caller must get exactly its own result, batches must really form, a failing batch must fail every one of its waiters
and nobody else, and TSAN must stay silent. The GPU counterpart (results equal direct calls) is tests/test_gpu_batcher.py."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_batcher_threads_under_tsan(tmp_path):
    cuda_inc, cuda_lib = "/usr/local/cuda/include", "/usr/local/cuda/lib64"
    if shutil.which("g++") is None or not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA headers")
    csrc = os.path.join(ROOT, "rag_era_b200", "csrc")
    exe = str(tmp_path / "batcher_tsan")
    r = subprocess.run(["g++", "-std=c++17", "-g", "-O1", "-fsanitize=thread", "-x", "c++", "-I" + cuda_inc, "-I" + os.path.join(ROOT, "include"),
                        "-I" + csrc, os.path.join(ROOT, "tests", "c", "batcher_tsan.cc"), os.path.join(csrc, "batcher.cu"), "-o", exe,
                        "-L" + cuda_lib, "-lcudart", "-Wl,-rpath," + cuda_lib, "-lpthread"], capture_output=True, text=True)
    if r.returncode != 0 and "sanitize" in r.stderr and "cannot find" in r.stderr:
        pytest.skip("libtsan is not installed")
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    if "FATAL: ThreadSanitizer" in r.stderr and "unexpected memory mapping" in r.stderr:
        pytest.skip("ThreadSanitizer cannot run in this sandbox (ASLR layout)")
    assert r.returncode == 0 and "result: OK" in r.stdout, (r.stdout[-800:], r.stderr[-1500:])
    assert "WARNING: ThreadSanitizer" not in r.stderr, r.stderr[-3000:]
    assert "largest=8" in r.stdout          # batches fill up to max_batch under load

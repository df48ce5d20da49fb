"""§8f N2 on CPU: the N-API addon's threading and handle lifetime under ThreadSanitizer, without Node and without a GPU.
integration/node/ragera_addon.cc (unmodified) + the mock Node-API host (tests/c/napi_mock.cc) + the REAL micro-batcher
(batcher.cu) are linked against tests/c/ragera_stub.cc, whose "searches" are fixed functions of the query and which aborts
on any call that reaches an index after rag_index_destroy. Checked: every Promise is fulfilled; `submit` answers through
the thread-safe function (no pool thread parked per request); requests beyond the batcher's slots fall back to pool
threads (RAG_ERR_BUSY) and are answered; destroyBatcher()/destroy() with calls still in flight answer them and only then
free the native objects; calls on a closed handle throw; finalizers free the wrappers; TSAN stays silent.
The numerics of the same flow are checked on a B200 by tests/test_gpu_napi.py."""
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

from test_gpu_napi import write_input

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_addon_lifetime_and_async_submit_under_tsan(tmp_path):
    cuda_inc, cuda_lib = "/usr/local/cuda/include", "/usr/local/cuda/lib64"
    if shutil.which("g++") is None or not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA headers")
    csrc = os.path.join(ROOT, "rag_era_b200", "csrc")
    exe = str(tmp_path / "napi_tsan")
    r = subprocess.run(["g++", "-std=c++17", "-g", "-O1", "-fsanitize=thread", "-x", "c++", "-I" + cuda_inc, "-I" + os.path.join(ROOT, "include"),
                        "-I" + csrc, "-I" + os.path.join(ROOT, "tests", "c", "node_api_stub"),
                        os.path.join(ROOT, "tests", "c", "napi_mock.cc"), os.path.join(ROOT, "integration", "node", "ragera_addon.cc"),
                        os.path.join(csrc, "batcher.cu"), os.path.join(ROOT, "tests", "c", "ragera_stub.cc"), "-o", exe,
                        "-L" + cuda_lib, "-lcudart", "-Wl,-rpath," + cuda_lib, "-lpthread"], capture_output=True, text=True)
    if r.returncode != 0 and "sanitize" in r.stderr and "cannot find" in r.stderr:
        pytest.skip("libtsan is not installed")
    assert r.returncode == 0, r.stderr[-2000:]
    n, d, B, k, kw_limit = 64, 8, 12, 10, 5
    Q = np.zeros((B, d), np.float32)
    Q[:, 0] = 101 + np.arange(B)                        # the stub's results are a function of the first element
    kw_counts = (np.arange(B) % (kw_limit + 1)).astype(np.uint32)
    kw_keys = (np.arange(B * kw_limit).reshape(B, kw_limit) + 7).astype(np.uint64)
    inp = str(tmp_path / "in.bin")
    write_input(inp, np.zeros((n, d), np.float32), Q, k, kw_limit, 0.3, np.zeros(n), np.arange(n), kw_keys, kw_counts)
    r = subprocess.run([exe, inp], capture_output=True, text=True, timeout=300)
    if "FATAL: ThreadSanitizer" in r.stderr and "unexpected memory mapping" in r.stderr:
        pytest.skip("ThreadSanitizer cannot run in this sandbox (ASLR layout)")
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-2000:])
    assert "WARNING: ThreadSanitizer" not in r.stderr and "STUB:" not in r.stderr, r.stderr[-3000:]
    out = json.loads(r.stdout)
    for b in range(B):
        tag = 101 + b
        g = out["single"][b]
        nres = 1 + tag % k
        assert g["keys"][:nres] == [tag * 1000 + i for i in range(nres)] and g["keys"][nres] == int(kw_keys[b, :kw_counts[b]].sum())
        assert g["vecIds"] == [tag] and g["usedRrf"] == (1 if kw_counts[b] else 0)
        assert out["in_flight"][b] == g and out["batched"][b] == g and out["one_call"][b]["keys"] == g["keys"]
        assert out["search"][b]["ids"] == [tag * 10 + i for i in range(k)] and out["doomed"][b] == out["search"][b]["ids"]
    assert out["via_tsfn"] == B                                                   # answered without a pool thread
    assert len(out["overflow"]) == 3 * B and all(out["overflow"][r] == out["single"][r % B] for r in range(3 * B))
    assert "destroyed" in out["closed_batcher"] and "destroyed" in out["closed_index"]
    assert "libragera error" in out["rejected"] and "Float32Array" in out["thrown"]

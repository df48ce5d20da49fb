#!/usr/bin/env python3
"""Small end-to-end pass over every kernel for compute-sanitizer (memcheck): stream (K1, K1m), exact (K1x),
fused K3+K4, K3/K4 throughput variants, K5 (all modes), tensor path (K2 pair), generator kernels."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle  # noqa: E402
import rag_era_b200 as rb  # noqa: E402
from rag_era_b200 import _native as N  # noqa: E402

n, d = 3000, 256
go = oracle.make_gen(n, n_clusters=8, dup_period=7, memory_rows=500)
gn = N.GenDesc.from_buffer_copy(bytes(go))
X = oracle.gen_rows(go, 0, n, d)
ok = True
with rb.VectorIndex(d, n, bf16_shadow=True) as idx:
    idx.generate(gn, n)
    Q = idx.generate_queries(gn, 0, 40)
    for path, B in ((N.PATH_STREAM, 1), (N.PATH_STREAM, 5), (N.PATH_EXACT, 2), (N.PATH_TENSOR, 40), (N.PATH_STREAM, 40)):
        r = idx.query(Q[:B], 10, path=path)
        for b in range(B):
            ei, es = oracle.topk(X, Q[b], 10)
            ok &= bool(np.array_equal(r.row(b)[0], ei) and np.array_equal(r.row(b)[1], es))
    f = idx.hybrid(Q[:3], rb.hybrid_opts(10, 4, 0.3, fresh_limit=5, now_ms=1_760_000_000_000), [[1, 2], [], [5, 6, 7, 8]])
    ok &= bool(f.counts.sum() > 0)
    m = idx.memory_retrieve(Q[:2], 5, 0.1, now_ms=1_760_000_000_000)
    fr = idx.freshness_scores([0.5, 0.9], [1, 2], [0, 1000], 5000)
    z = idx.rrf_fuse([[1, 2, 3]], [[2, 4]])
    ok &= bool(z.counts[0] == 4)
# the fp16 shadow of the normalised rows (pre-normalised epilogue), a bf16 corpus (bf16 x bf16 + multi-query bf16 stream
# kernel), and the small-batch call replayed as a CUDA graph (second call = replay)
with rb.VectorIndex(d, n, shadow="f16") as idx:
    idx.generate(gn, n)
    r = idx.query(Q[:40], 10, path=N.PATH_TENSOR)
    for b in range(0, 40, 7):
        ei, es = oracle.topk(X, Q[b], 10)
        ok &= bool(np.array_equal(r.row(b)[0], ei) and np.array_equal(r.row(b)[1], es))
    for _ in range(2):
        f = idx.hybrid(Q[:1], rb.hybrid_opts(10, 4, 0.3, path=N.PATH_STREAM), [[1, 2, 3]])
        e = oracle.hybrid_search(X, Q[0], 10, 0.3, [1, 2, 3])
        ok &= bool(np.array_equal(f.row(0)["keys"], e["keys"]) and np.array_equal(f.row(0)["scores"], e["scores"]))
Xb = oracle.gen_rows(go, 0, n, d, dtype=oracle.BF16)
with rb.VectorIndex(d, n, dtype=N.BF16) as idx:
    idx.generate(gn, n)
    for path, B in ((N.PATH_TENSOR, 40), (N.PATH_STREAM, 6)):
        r = idx.query(Q[:B], 10, path=path)
        for b in range(0, B, 5):
            ei, es = oracle.topk(Xb, Q[b], 10)
            ok &= bool(np.array_equal(r.row(b)[0], ei) and np.array_equal(r.row(b)[1], es))
print("sanitize_small:", "OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)

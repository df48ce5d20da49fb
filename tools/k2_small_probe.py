#!/usr/bin/env python3
"""Small-batch crossover on C2-sized data: stream (K1/K1m) vs tensor (K2 tf32 on fp32 rows / bf16 shadow)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rag_era_b200 as rb
from rag_era_b200 import _native as N

rows, d = 1_000_000, 1536
gen = N.GenDesc(0xC0FFEE, 0xBEEF, 0xF00D, rows, 4096, 0.6, 0.5, 0, 0, 0)
for shadow in (False, True):
    with rb.VectorIndex(d, rows, bf16_shadow=shadow) as idx:
        idx.generate(gen, rows)
        Q = idx.generate_queries(gen, 0, 256)
        idx.profile_enable(True)
        for B in (4, 8, 16, 32, 64, 128, 256):
            out = []
            for path, cls in ((N.PATH_STREAM, "stream"), (N.PATH_TENSOR, "tensor")):
                if path == N.PATH_STREAM and B > 32:
                    out.append("stream   -   ")
                    continue
                for _ in range(2):
                    idx.query(Q[:B], 10, path=path)
                idx.profile_read()
                for _ in range(5):
                    r = idx.query(Q[:B], 10, path=path, flags=N.SEARCH_NO_ESCALATE)
                p = idx.profile_read()
                out.append(f"{cls} {p[cls][0] / 5:7.3f} ms (cert {int(r.certified.sum())}/{B})")
            print(f"shadow={shadow} B={B:3d}: " + " | ".join(out))

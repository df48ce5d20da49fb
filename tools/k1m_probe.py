#!/usr/bin/env python3
"""Time the stream path at small batches (K1 / K1m) on C2-sized data: ms per pass and per query."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rag_era_b200 as rb
from rag_era_b200 import _native as N

rows, d = 1_000_000, 1536
gen = N.GenDesc(0xC0FFEE, 0xBEEF, 0xF00D, rows, 4096, 0.6, 0.5, 0, 0, 0)
for dt, name in ((N.F32, "f32"), (N.BF16, "bf16")):
    with rb.VectorIndex(d, rows, dtype=dt) as idx:
        idx.generate(gen, rows)
        Q = idx.generate_queries(gen, 0, 64)
        idx.profile_enable(True)
        for B in (1, 2, 4, 8, 12, 16, 32):
            for _ in range(3):
                idx.query(Q[:B], 10, path=N.PATH_STREAM)
            idx.profile_read()
            n = 10
            for _ in range(n):
                idx.query(Q[:B], 10, path=N.PATH_STREAM)
            p = idx.profile_read()
            ms = p["stream"][0] / n
            gb = rows * d * (4 if dt == N.F32 else 2) / 1e9
            print(f"{name} B={B:3d}: stream {ms:7.3f} ms/batch  {ms / B:7.4f} ms/query  passes-equivalent BW {gb / (ms * 1e-3) / 1e3:6.2f} TB/s x {B} = per-query-effective {B * gb / (ms * 1e-3) / 1e3:7.2f} TB/s  tail {sum(v[0] for k, v in p.items() if k != 'stream') / n:6.3f} ms")

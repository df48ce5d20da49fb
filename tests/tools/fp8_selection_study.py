#!/usr/bin/env python3
"""Design study for DESIGN §8 item 2 (CPU only, numpy): could an fp8-e4m3 copy of the corpus SELECT candidates for
the batched tensor path (K4 would still rescore exactly)? Measures, on the synthetic corpus of include/ragera_gen.h,
  (1) the selection error |cos_lowprec - cos_exact| for bf16 and for e4m3 operands (per-row scale = max|x|/448),
  (2) the gap between the k-th and the K'-th exact cosine per query,
and from both the fraction of queries the certification rule (exact k-th > approx K'-th + eps) would pass.

    python tests/tools/fp8_selection_study.py [rows] [queries]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402  (a study script, not product code)


def round_bf16(a):
    u = a.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def round_e4m3(a):
    """Round-to-nearest-even to OCP e4m3 (3 mantissa bits, max 448, subnormals below 2^-6). Checked against the host
    implementation of CUDA's own `__nv_cvt_float_to_fp8(x, __NV_SATFINITE, __NV_E4M3)` (cuda_fp8.h compiled with g++) on a
    sweep of 4000 values across the normal and subnormal ranges: identical."""
    a = a.astype(np.float64)
    s = np.sign(a)
    m = np.abs(a)
    m = np.minimum(m, 448.0)
    e = np.floor(np.log2(np.maximum(m, 2.0 ** -20)))
    e = np.maximum(e, -6.0)                     # subnormal range shares the exponent of 2^-6
    q = 2.0 ** (e - 3)                          # spacing: 3 mantissa bits
    return (s * np.round(m / q) * q).astype(np.float32)


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    nq = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    d, k = 1536, 10
    g = oracle.make_gen(rows, n_clusters=max(16, rows * 4096 // 1_000_000))   # the density of C2 (244 rows per cluster)
    X = oracle.gen_rows(g, 0, rows, d)
    Q = oracle.gen_queries(g, 0, nq, d)
    Xn = X / np.linalg.norm(X.astype(np.float64), axis=1, keepdims=True).astype(np.float32)
    Qn = Q / np.linalg.norm(Q.astype(np.float64), axis=1, keepdims=True).astype(np.float32)
    exact = (Qn.astype(np.float64) @ Xn.astype(np.float64).T)                  # [nq][rows]

    def lowprec(rx, rq, per_row_scale):
        if per_row_scale:
            sx = np.max(np.abs(X), axis=1, keepdims=True) / 448.0
            sq = np.max(np.abs(Q), axis=1, keepdims=True) / 448.0
            Xl, Ql = rx(X / sx) * sx, rq(Q / sq) * sq
        else:
            Xl, Ql = rx(X), rq(Q)
        nx = np.linalg.norm(Xl.astype(np.float64), axis=1)
        nqv = np.linalg.norm(Ql.astype(np.float64), axis=1)
        return (Ql.astype(np.float64) @ Xl.astype(np.float64).T) / nqv[:, None] / nx[None, :]

    srt = -np.sort(-exact, axis=1)
    print(f"corpus {rows} x {d}, {nq} queries, k={k}; exact top-1 {srt[:, 0].mean():.3f}, k-th {srt[:, k - 1].mean():.3f}")
    for name, approx in (("bf16", lowprec(round_bf16, round_bf16, False)), ("e4m3 (per-row scale)", lowprec(round_e4m3, round_e4m3, True)),
                         ("e4m3 corpus x bf16 query", lowprec(round_e4m3, round_bf16, True))):
        err = np.abs(approx - exact)
        # the rows that matter: anything within 0.1 of the k-th score
        near = exact > (srt[:, k - 1:k] - 0.1)
        sig, mx = err[near].std(), err[near].max()
        print(f"{name:26s} sigma {sig:.2e}  max {mx:.2e}  (over {near.sum()} near-top pairs; all pairs max {err.max():.2e})")
        for kp in (32, 48, 64, 128):
            for nsig in (6, 11):
                eps = nsig * sig
                a_srt = -np.sort(-approx, axis=1)
                t = a_srt[:, kp - 1]                                     # approx score of the K'-th candidate
                cand = approx >= t[:, None]
                kth_exact_in_cand = -np.sort(-np.where(cand, exact, -1.0), axis=1)[:, k - 1]
                ok_cert = kth_exact_in_cand > t + eps                    # K4's certification rule
                true_topk = np.argsort(-exact, axis=1, kind="stable")[:, :k]
                complete = np.array([cand[b, true_topk[b]].all() for b in range(nq)])
                print(f"    K'={kp:3d} eps={nsig:2d}sigma={eps:.1e}: certified {ok_cert.mean() * 100:5.1f}%  (true top-{k} inside the candidates {complete.mean() * 100:5.1f}%)")


if __name__ == "__main__":
    main()

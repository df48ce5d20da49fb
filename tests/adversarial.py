"""Adversarial corpus for the tensor path's certification (test infrastructure).

Real embedding models have a few outlier dimensions that carry most of a vector's energy. When the values in
those dimensions sit just below a bf16 rounding midpoint, the rounding errors of all of them have the same sign
and the bf16 copy of the row is SHORTER than the row by ~2^-8 relative — an order of magnitude more than the
~11 sigma statistical bound of round 1 (0.024/sqrt(ld)), which assumed independent errors.

The construction (fp32 index + BF16 shadow; the query goes in as bf16 of q/||q||):
  q      outlier dims = +-8*(1+0.98/256) (bf16 rounds them DOWN to +-8), the rest small bf16-exact values
  star   row `star` = q itself: exact cosine 1.0, the true top-1; its bf16 copy scores (1-delta) ~ 0.9962
  decoys ~80 bf16-exact rows bf16(q) + p_j with exact cosines 0.9970..0.9990: their copies are exact, so all of
         them outrank `star` on the tensor path and push it out of the K'=48 candidates
With the statistical bound the query certifies although `star` is missing from the answer; the rigorous bound
(measured residuals rho_q, rho_x) refuses, the query escalates to the fp32 stream path and the answer is exact.
"""
import numpy as np

import oracle


def bf16_round(a):
    return oracle.bf16_to_f32(oracle.f32_to_bf16(a))


def f16_round(a):
    return np.asarray(a, dtype=np.float32).astype(np.float16).astype(np.float32)


def build(n=6000, d=1536, n_out=16, n_decoy=80, seed=7):
    rng = np.random.default_rng(seed)
    q = bf16_round((0.01 * rng.standard_normal(d)).astype(np.float32))
    sign = rng.choice([-1.0, 1.0], n_out).astype(np.float32)
    q[:n_out] = sign * np.float32(8.0 * (1 + 0.98 / 256))
    qt = bf16_round(q)
    assert np.all(np.abs(qt[:n_out]) == 8.0)
    X = (0.3 * rng.standard_normal((n, d))).astype(np.float32)
    star = n - 5
    X[star] = q
    nq = np.linalg.norm(q.astype(np.float64))
    rows = rng.choice(np.arange(10, n - 10), n_decoy, replace=False)
    rows = rows[rows != star]
    for j, r in enumerate(rows):
        target = 0.9970 + 0.0020 * j / len(rows)          # exact cosine of this decoy
        p = rng.standard_normal(d).astype(np.float32)
        p[:n_out] = 0
        p *= np.float32(nq * np.sqrt(2 * (1 - target)) / np.linalg.norm(p))
        X[r] = bf16_round(qt + p)
    return X, q, star, rows


def operands(X, q, rows="bf16"):
    """The tensor path's operands in fp64 (gen.cu's table): (q~, x~ in cosine units, rho_q, rho_x).
    rows: "bf16" = bf16 shadow of fp32 rows, "f16" = fp16 shadow of the normalised rows, "exact" = a bf16 corpus.
    The query operand has the rows' format (kind::f16 takes one format for both): fp16 with the fp16 shadow, else bf16."""
    Xd, qd = X.astype(np.float64), q.astype(np.float64)
    nx, nq = np.linalg.norm(Xd, axis=1, keepdims=True), np.linalg.norm(qd)
    uq = qd / nq
    qt = (f16_round(uq) if rows == "f16" else bf16_round(uq.astype(np.float32))).astype(np.float64)
    ux = Xd / nx
    if rows == "bf16":
        xt = bf16_round(X).astype(np.float64) / nx
    elif rows == "f16":
        xt = f16_round(ux).astype(np.float64)
    else:
        xt = ux
    return qt, xt, np.linalg.norm(qt - uq), np.linalg.norm(xt - ux, axis=1).max()


def emulate(X, q, rows="bf16"):
    """(exact cosines, the tensor path's scores, rho_q, rho_x) in fp64 emulation."""
    Xd, qd = X.astype(np.float64), q.astype(np.float64)
    exact = (Xd @ qd) / np.linalg.norm(Xd, axis=1) / np.linalg.norm(qd)
    qt, xt, rho_q, rho_x = operands(X, q, rows)
    return exact, xt @ qt, rho_q, rho_x


def emulate_shadow_stream(X, q):
    """(exact cosines, the shadow stream path's scores in cosine units, rho_x): the fp32 query as it is against the fp16
    shadow of the normalised rows — q.x~ / ||q||, in fp64 (the fp32 accumulation is bounded separately)."""
    Xd, qd = X.astype(np.float64), q.astype(np.float64)
    nx, nq = np.linalg.norm(Xd, axis=1, keepdims=True), np.linalg.norm(qd)
    ux = Xd / nx
    xt = f16_round(ux).astype(np.float64)
    return (ux @ qd) / nq, (xt @ qd) / nq, np.linalg.norm(xt - ux, axis=1).max()

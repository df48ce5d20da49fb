#!/usr/bin/env python3
"""bench.py — hybrid-search throughput/latency of the retrieval hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c2b|c1|c4|c4f|c5|c3s|c1s] [--impl reference]

One "step" = one hybrid search (cosine scoring of every chunk → top-k → min-cosine filter →
RRF with the keyword list) of one query batch over the whole corpus.

  value     queries/s with the batch already resident in HBM (device-timed with CUDA events on
            the library stream, K steps back to back between barriers, max over ranks)
  e2e       the same metric through the public call rag_hybrid_search with HOST buffers
            (H2D of queries+keyword lists and D2H of results inside the timed region)
  roofline  the dominant kernel (K1 stream scoring) against the measured HBM copy bandwidth
  cpu_baseline  the oracle (a port of the reference's single-threaded JS algorithm) on host cores

Workloads (BASELINE.json configs): c3 = 10M x 1536 fp32 batch 1 (default; the north-star
target), c2 = 1M x 1536 fp32 batch 1 (c2b: batch 1024, tcgen05 path), c1 = 10k x 1536 fp32 batch 1
(L2-resident), c4 = 2M memory + 5M doc rows, three lists, batch 256, c5 = 50M x 1536 bf16 batch 1024.
c3s / c1s = c3 / c1 with the single query scored by streaming the fp16 shadow of the normalised rows (an EXTRA mode:
half the bytes, same certified-exact results; the BASELINE config is the fp32 stream).
The default line carries c2 / c2b / c1 / c4 (one GPU) and c5, c3s (every N) under `extra`.
For N > 1 the SAME corpus is row-sharded over the ranks (strong scaling): every rank scores
its shard, the ranks' exact local top-k lists are exchanged inside the final fusion kernel through
peer-to-peer mailboxes over NVLink (RAGERA_COMM=nccl: one ncclAllGather instead), and the merge +
fusion runs on every rank. torch.distributed (NCCL) only carries the bench's own barriers/reductions.

`--impl reference` times the reference's CPU algorithm (oracle port) for the same metric: ONE host
thread, because the reference is single-threaded JavaScript, on a bounded row sample
extrapolated to the full corpus (scoring linearly, the full stable sort as n log n); the all-threads
figure is a labelled courtesy field.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: rows, dim, dtype, batch, vectorTopK, keywordLimit, minVectorScore, show
    "c1": dict(rows=10_000, dim=1536, dtype="f32", batch=1, vector_top_k=5, keyword_limit=5, min_score=0.3, show=3,
               desc="C1 search_knowledge: 10k x 1536 fp32, vectorTopK=5 keywordLimit=5 RRF(k=60) top-3, batch 1"),
    "c2": dict(rows=1_000_000, dim=1536, dtype="f32", batch=1, vector_top_k=10, keyword_limit=10, min_score=0.3, show=8,
               desc="C2 deep_search: 1M x 1536 fp32, vectorTopK=10 keywordLimit=10 RRF(k=60) top-8, batch 1"),
    "c3": dict(rows=10_000_000, dim=1536, dtype="f32", batch=1, vector_top_k=10, keyword_limit=10, min_score=0.3, show=8,
               desc="C3: 10M x 1536 fp32, vectorTopK=10 keywordLimit=10 RRF(k=60) top-8, batch 1"),
    "c3h": dict(rows=10_000_000, dim=1536, dtype="bf16", batch=1, vector_top_k=10, keyword_limit=10, min_score=0.3, show=8,
                desc="C3 with a bf16 corpus: 10M x 1536 bf16, vectorTopK=10 keywordLimit=10 RRF(k=60) top-8, batch 1 (stream path)"),
    "c3s": dict(rows=10_000_000, dim=1536, dtype="f32", batch=1, vector_top_k=10, keyword_limit=10, min_score=0.3, show=8,
                path="shadow_stream", shadow="f16",
                desc="C3 on the fp16 shadow (EXTRA mode, not the BASELINE config): 10M x 1536 fp32 (+fp16 shadow of the normalised rows), "
                     "batch 1 scored by streaming the shadow (2 B/element), K'=32, certified by the rows' measured rounding residual; "
                     "vectorTopK=10 keywordLimit=10 RRF(k=60) top-8"),
    "c1s": dict(rows=10_000, dim=1536, dtype="f32", batch=1, vector_top_k=5, keyword_limit=5, min_score=0.3, show=3,
                path="shadow_stream", shadow="f16",
                desc="C1 on the fp16 shadow (EXTRA mode): 10k x 1536 fp32 (+fp16 shadow), batch 1 scored by streaming the shadow; "
                     "vectorTopK=5 keywordLimit=5 RRF(k=60) top-3"),
    "c2b": dict(rows=1_000_000, dim=1536, dtype="f32", batch=1024, vector_top_k=10, keyword_limit=10, min_score=0.3, show=8,
                path="tensor", shadow="f16",
                desc="C2 deep_search batched: 1M x 1536 fp32 (+fp16 shadow of the normalised rows), vectorTopK=10 keywordLimit=10 RRF(k=60) top-8, batch 1024, tcgen05 kind::f16 path"),
    "c2bb": dict(rows=1_000_000, dim=1536, dtype="f32", batch=1024, vector_top_k=10, keyword_limit=10, min_score=0.3, show=8,
                 path="tensor", shadow="bf16",
                 desc="C2 batched with a BF16 shadow: 1M x 1536 fp32 (+bf16 shadow), batch 1024, tcgen05 path (wider candidate window: bf16's rounding bound)"),
    "c2br": dict(rows=1_000_000, dim=1536, dtype="f32", batch=1024, vector_top_k=10, keyword_limit=10, min_score=0.3, show=8,
                 path="tensor", shadow="f16", replicas=True,
                 desc="C2 batched, REPLICAS mode: the whole 1M x 1536 fp32 corpus (+fp16 shadow) on every GPU, the 1024-query batch split over "
                      "the ranks, no exchange — the alternative to row shards for corpora that do not fill the box (SURVEY §8e)"),
    "c4f": dict(rows=7_000_000, dim=1536, dtype="f32", batch=256, vector_top_k=10, keyword_limit=10, min_score=0.3, show=8,
                path="tensor", shadow=None, memory_rows=2_000_000, fresh_limit=10,
                desc="C4 memory+RAG unified, fp32 operand: 2M memory + 5M doc rows x 1536 fp32 scored as tf32 (no shadow), vector+keyword+freshness lists fused by RRF, batch 256, tcgen05 kind::tf32"),
    "c4": dict(rows=7_000_000, dim=1536, dtype="f32", batch=256, vector_top_k=10, keyword_limit=10, min_score=0.3, show=8,
               path="tensor", shadow="f16", memory_rows=2_000_000, fresh_limit=10,
               desc="C4 memory+RAG unified: 2M memory + 5M doc rows x 1536 fp32 (+fp16 shadow of the normalised rows), vector+keyword+freshness lists fused by RRF, batch 256, tcgen05 kind::f16 path"),
    "c5": dict(rows=50_000_000, dim=1536, dtype="bf16", batch=1024, vector_top_k=10, keyword_limit=10, min_score=0.3, show=8,
               path="tensor",
               desc="C5: 50M x 1536 bf16 row-sharded, vectorTopK=10 keywordLimit=10 RRF(k=60) top-8, batch 1024, tcgen05 path, "
                    "local top-k exchanged through peer-to-peer mailboxes inside the fusion kernel"),
}
SEEDS = dict(seed=0xC0FFEE, query_seed=0xBEEF, meta_seed=0xF00D)
KW_SEED = 0xFACE
METRIC, UNIT = "hybrid_search_qps", "queries/s"


def peaks():
    """(hbm GB/s, bf16 TFLOP/s sustained, bf16 TFLOP/s burst, source)"""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return (float(j["hbm_gbs"]), float(j.get("bf16_tflops_sustained", j["bf16_tflops"])), float(j["bf16_tflops"]),
                    "measured (MEASURED_PEAKS.json)")
        except Exception:
            pass
    return 6650.0, 1400.0, 1590.0, "fallback (B200_PROFILING.md: 6.65 TB/s, 1.59 PFLOP/s burst / ~1.4 sustained)"


def traffic_ratio(kernel: str, name: str):
    """DRAM traffic / algorithmic bytes of the dominant kernel from the committed ncu --set full capture
    (profiles/traffic.json); None if this workload has not been captured."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[kernel]
        e = t.get(name) or t.get({"c3": "c2", "c2": "c3"}.get(name, ""))   # same kernel, same access pattern
        return float(e.get("ratio", e.get("traffic_over_algorithmic")))
    except Exception:
        return None


def keyword_lists(top_ids: np.ndarray, rows: int, kl: int, rng: np.random.Generator):
    """ceil(30%) of each list drawn from the true vector top-k (exercises 'both'),
    the rest uniform random rows, in a seeded order. Key map = identity."""
    out = []
    for b in range(top_ids.shape[0]):
        n_hit = min(kl, -(-3 * kl // 10))
        hits = [int(x) for x in top_ids[b, :n_hit]]
        hits += [int(x) for x in rng.integers(0, rows, kl - len(hits))]
        rng.shuffle(hits)
        out.append(hits)
    return out


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons of one GPU sampled during the timed region (NVML)."""

    def __init__(self, device: int):
        super().__init__(daemon=True)
        self.device, self.samples, self.reasons, self._stop_evt = device, [], set(), threading.Event()
        self.max_mhz = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.01)

    def stop(self):
        self._stop_evt.set()
        self.join(2)
        s = sorted(self.samples)
        return dict(sm_mhz=(s[len(s) // 2] if s else None), sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons),
                    samples=len(s))


# --------------------------------------------------------------------------------------------
# CPU legs (the only places bench.py may execute oracle/)
# --------------------------------------------------------------------------------------------
def cpu_reference_leg(w, steps, warmup, threads, sample_rows):
    """The reference's algorithm (oracle port: fp64 left-to-right cosine with the norms recomputed for every row,
    score ALL rows, full stable sort, slice k, min-cosine filter, RRF) over the first `sample_rows` rows of
    workload `w`, `threads` host threads (1 = what the reference is). The per-query time is extrapolated to the
    full corpus: the scoring part linearly (it is exactly linear in rows), the full stable sort as n*log2(n).
    Returns (qps at full size, seconds per query at full size, description dict)."""
    import oracle

    rows, d = w["rows"], w["dim"]
    s_rows = min(rows, sample_rows)
    g = oracle.make_gen(rows, **SEEDS)
    X = oracle.gen_rows(g, 0, s_rows, d, threads=os.cpu_count() or 1)
    Q = oracle.gen_queries(g, 0, steps + warmup, d)
    rng = np.random.default_rng(KW_SEED)
    cfg = oracle.RRFConfig()
    full, score_only = [], []
    for i in range(steps + warmup):
        kw = [int(x) for x in rng.integers(0, s_rows, w["keyword_limit"])]
        t0 = time.perf_counter()
        ids, sc = oracle.topk(X, Q[i], w["vector_top_k"], faithful_sort=True, threads=threads)
        ids, sc = oracle.filter_min_score(ids, sc, w["min_score"])
        oracle.rrf(ids, kw, cfg)
        dt = time.perf_counter() - t0
        if i >= warmup:
            full.append(dt)
    # the same scan with a partial selection instead of the full sort: the difference is the sort's share
    for i in range(min(2, steps)):
        t0 = time.perf_counter()
        oracle.topk(X, Q[warmup + i], w["vector_top_k"], faithful_sort=False, threads=threads)
        score_only.append(time.perf_counter() - t0)
    t_full, t_score = float(np.mean(full)), float(np.mean(score_only))
    t_sort = max(0.0, t_full - t_score)
    scale = rows / s_rows
    sort_scale = scale * (np.log2(rows) / np.log2(max(s_rows, 2)))
    per_query = (t_full - t_sort) * scale + t_sort * sort_scale
    info = {"sample_rows": int(s_rows), "rows": int(rows), "queries_timed": len(full), "threads": int(threads),
            "seconds_per_sample_query": t_full, "sort_share_of_sample_query": t_sort / t_full if t_full else 0.0,
            "scale_scoring": scale, "scale_sort_nlogn": float(sort_scale),
            "full_size_query_timed": bool(s_rows == rows)}
    sample = (f"{len(full)} queries x first {s_rows} of {rows} rows on {threads} host thread(s): fp64 left-to-right cosine, norms "
              f"recomputed per row, score-all + full stable sort, filter, RRF; scoring time x{scale:g} (linear in rows), "
              f"sort time x{sort_scale:.3g} (n log n)" + ("" if s_rows == rows else "; no full-size CPU query is timed"))
    return 1.0 / per_query, per_query, sample, info


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    all_threads = os.cpu_count() or 1
    steps, warmup = args.steps or 5, args.warmup if args.warmup is not None else 1
    t0 = time.perf_counter()
    # Variant A: ONE thread — the reference is single-threaded JavaScript. Bound the run to
    # about two minutes: calibrate on 20k rows, then size the row sample for steps+warmup queries (+2 sort-free ones).
    _, cal, _, _ = cpu_reference_leg(dict(w, rows=20_000), 1, 0, 1, 20_000)
    budget_rows = int(20_000 * 75.0 / max(cal * (steps + warmup + 2), 1e-9))
    sample_rows = max(20_000, min(w["rows"], 1_000_000, budget_rows))
    qps, per_query, sample, info = cpu_reference_leg(w, steps, warmup, 1, sample_rows)
    # Variant B, courtesy only: the same port row-parallel over every host thread (NOT what the reference does)
    qps_mt, _, _, info_mt = cpu_reference_leg(w, min(steps, 3), 1, all_threads, sample_rows)
    wall = time.perf_counter() - t0
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": per_query * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config_of(w, args.workload, args.gpus),
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "extrapolated": info,
        "all_threads_courtesy": {"value": qps_mt, "unit": UNIT, "cores": all_threads,
                                 "note": "row-parallel over all host threads — NOT the reference's algorithm (single-threaded JS); "
                                         "changes with the box, never used as the baseline"},
        "note": "reference is TypeScript (no Node here; its dense arithmetic lives in un-vendored llamaindex@0.12.1): this is the "
                "oracle port on ONE host thread. value and ms_per_step are the EXTRAPOLATED full-corpus figures (see "
                "`extrapolated`), so steps*ms_per_step exceeds this run's wall time by design: each step scans only the sample.",
        "wall_s": wall, "wall_s_per_step": info["seconds_per_sample_query"],
    }
    print(json.dumps(line), flush=True)


def config_of(w, name, gpus):
    return {"workload": w["desc"], "name": name, "rows": w["rows"], "dim": w["dim"], "corpus_dtype": w["dtype"],
            "batch": w["batch"], "vector_top_k": w["vector_top_k"], "keyword_limit": w["keyword_limit"],
            "min_vector_score": w["min_score"], "rrf": {"k": 60, "vector_weight": 1.0, "keyword_weight": 1.0, "both_bonus": 0.1},
            "display_top": w["show"], "sharding": f"rows/{gpus}" if gpus > 1 else "none",
            "cache": "corpus streamed per step is larger than L2 (126 MB); no flush needed" if w["rows"] * w["dim"] * 4 > 2.5e8
            else "corpus is L2-resident (latency-bound case); reported as such",
            "baseline_config": {"c1": "BASELINE.json configs[0]", "c2": "configs[1] (batch 1)", "c2b": "configs[1] (batch 1024)",
                                "c3": "configs[2] — the north_star target (>= 80% of HBM roofline, batch-1 top-k over 10M x 1536 fp32 on one "
                                      "B200) and the default line: its 10M rows still give every GPU real work when the same corpus is "
                                      "sharded over 8; configs[1] is reported in the same line under extra.c2 / extra.c2b",
                                "c4": "configs[3]", "c4f": "configs[3], fp32 operand", "c5": "configs[4]"}.get(name, name)}


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def measure_workload(rb, N, w, name, steps, warmup, dist, rank, world, device, do_cpu):
    from rag_era_b200.sharded import create_sharded_index, leave_exchange, shard_range

    rows, d, B = w["rows"], w["dim"], w["batch"]
    # REPLICAS mode (SURVEY §8e's alternative for corpora that do not fill the box): every rank holds the WHOLE corpus and
    # answers its own 1/N of the batch; no exchange at all. Below, such a rank behaves like a single-GPU run of B/N queries,
    # only the clock (barriers, max over ranks) and the final sum are shared.
    replicas = world if (w.get("replicas") and world > 1) else 1
    if replicas > 1:
        B = max(1, B // replicas)
    sharded = world > 1 and replicas == 1
    dt = N.F32 if w["dtype"] == "f32" else N.BF16
    now_ms = 1_760_000_000_000
    gen = N.GenDesc(SEEDS["seed"], SEEDS["query_seed"], SEEDS["meta_seed"], rows, 4096, 0.6, 0.5, 0,
                    w.get("memory_rows", 0), now_ms)
    tensor = w.get("path") == "tensor"
    on_shadow = w.get("path") == "shadow_stream"
    path = N.PATH_TENSOR if tensor else (N.PATH_SHADOW_STREAM if on_shadow else N.PATH_STREAM)
    shadow = w.get("shadow") if dt == N.F32 else None
    shard_rows = None
    if sharded:
        idx = create_sharded_index(dist, rows, d, dt, device, shadow=shadow, max_batch=max(B, 32), max_k=w["vector_top_k"])
        base, n_local = shard_range(rows, world, rank)
    else:
        idx = rb.VectorIndex(d, rows, dtype=dt, device=device, shadow=shadow)
        base, n_local = 0, rows
    idx.generate(gen, n_local)
    if sharded and tensor and os.environ.get("RAGERA_BENCH_BALANCE", "0") != "0":
        # OFF by default (measured: it does not help). The batch ends with the SLOWEST shard, and the GPUs of one box differ by
        # several percent under the power cap (round 1: up to 15% at N=8). Calibrate: a few batches on the even split, every
        # rank's own scoring-kernel time, then contiguous shards sized by measured speed (boundaries and id_base move, nothing
        # else). Result on 4 GPUs (C5): the rank that calibrated fastest (30.9 ms vs 32.7-32.9) got 6% more rows and then ran
        # 13% SLOWER than the others (34.9 vs 31.0-31.9 ms) — a GPU's speed under the power cap is a thermal state that moves
        # within seconds, not a property a short calibration can measure. profiles/r02_scaling.md.
        from rag_era_b200.sharded import balanced_ranges

        Qc = idx.generate_queries(gen, 0, B)
        oc = rb.hybrid_opts(w["vector_top_k"], 0, w["min_score"], path=path)
        idx.stage_batch(Qc, [[] for _ in range(B)], 0)
        for _ in range(2):
            idx.hybrid_staged(B, oc)
        idx.sync()
        dist.barrier()
        idx.profile_enable(True)
        idx.profile_read()
        for _ in range(4):
            idx.hybrid_staged(B, oc)
        idx.sync()
        k_ms = idx.profile_read()["tensor"][0] / 4
        idx.profile_enable(False)
        per = [None] * world
        dist.all_gather_object(per, (n_local, k_ms))
        ranges = balanced_ranges(rows, [p[0] for p in per], [p[1] for p in per])
        leave_exchange(dist, idx)
        idx.close()
        base, n_local = ranges[rank]
        idx = create_sharded_index(dist, rows, d, dt, device, shadow=shadow, max_batch=max(B, 32), max_k=w["vector_top_k"], rows=ranges[rank])
        idx.generate(gen, n_local)
        shard_rows = {"rows": [r[1] for r in ranges], "calibration_kernel_ms": [round(p[1], 3) for p in per],
                      "note": "contiguous shards sized by measured per-rank scoring speed (untimed calibration on the even split)"}
    total = steps + warmup
    n_pool = min(total, 8) if B >= 64 else total       # big batches cycle through a pool of distinct batches
    Q = idx.generate_queries(gen, (rank * n_pool * B) if replicas > 1 else 0, n_pool * B)   # replicas: each rank its own queries
    o = rb.hybrid_opts(w["vector_top_k"], w["keyword_limit"], w["min_score"], path=path, slack=int(os.environ.get("RAGERA_BENCH_SLACK", "0")),
                       fresh_limit=w.get("fresh_limit", 0), fresh_weight=1.0, now_ms=now_ms)

    # setup (untimed): true vector top-k of every query → keyword lists with ~30% overlap
    # (in chunks of one batch — at least 32 queries — so that a sharded index's mailboxes, sized for the batch, hold them)
    chunk = max(B, 32)
    tops = [idx.query(Q[lo:lo + chunk], w["vector_top_k"], path=path) for lo in range(0, len(Q), chunk)]
    top_ids = np.vstack([t.ids for t in tops])
    kw = keyword_lists(top_ids, rows, w["keyword_limit"], np.random.default_rng(KW_SEED))
    certified_setup = int(sum(int(t.certified.sum()) for t in tops))
    first_pass = None
    if tensor:
        # first-pass certification of ONE batch by the tensor path alone (no escalation), under both bounds:
        # the default rigorous one (a proof) and the round-1 statistical one
        rig = idx.query(Q[:B], w["vector_top_k"], path=path, flags=N.SEARCH_NO_ESCALATE)
        stat = idx.query(Q[:B], w["vector_top_k"], path=path, flags=N.SEARCH_NO_ESCALATE | N.SEARCH_STAT_EPS)
        first_pass = {"queries": B, "rigorous_bound": int(rig.certified.sum()), "statistical_bound": int(stat.certified.sum()),
                      "row_residual_rho_x": idx.row_residual()}
    elif on_shadow:
        nq = min(len(Q), 32)
        rig = idx.query(Q[:nq], w["vector_top_k"], path=path, flags=N.SEARCH_NO_ESCALATE)
        first_pass = {"queries": nq, "rigorous_bound": int(rig.certified.sum()), "row_residual_rho_x": idx.row_residual(),
                      "note": "queries certified by the shadow pass alone; the rest re-run on the fp32 stream inside the product call"}

    def barrier():
        idx.sync()
        if world > 1:
            dist.barrier()

    # ---- device-timed: batch resident in HBM ------------------------------------------------
    idx.stage_batch(Q, kw, w["keyword_limit"])
    for i in range(warmup):
        idx.stage_window((i % n_pool) * B, B)
        idx.hybrid_staged(B, o)
    barrier()
    idx.profile_enable(True)
    idx.profile_read()
    sampler = ClockSampler(device)
    sampler.start()
    l0 = idx.launch_count
    idx.certified_totals()      # reset the device-side counters
    barrier()
    idx.timer_start()
    for i in range(warmup, total):
        idx.stage_window((i % n_pool) * B, B)
        idx.hybrid_staged(B, o)
    ms = idx.timer_stop()
    barrier()
    clocks = sampler.stop()
    launches = idx.launch_count - l0
    prof = idx.profile_read()
    idx.profile_enable(False)
    timed_certified, timed_queries = idx.certified_totals()   # every query of every timed step, counted on the device
    last = idx.fetch_fused(B, o)
    parity = None
    if rank == 0 and os.environ.get("RAGERA_BENCH_PARITY", "1") != "0":
        parity = parity_check(w, gen, Q[((total - 1) % n_pool) * B:][:B], last, d, dt == N.BF16)   # untimed; the oracle as the checker
    if world > 1:
        import torch

        t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{device}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())

    # ---- end to end: host buffers through rag_hybrid_search ---------------------------------
    # inputs live in pinned host memory (what the N-API shim backs its typed arrays with); every call
    # copies them H2D, runs the pipeline, copies the packed result block D2H and unpacks it
    kl = w["keyword_limit"]
    Qp = idx.pinned_array((n_pool * B, d), np.float32)
    Qp[:] = Q
    kwk = np.zeros((n_pool * B, max(kl, 1)), dtype=np.uint64)
    kwc = np.zeros(n_pool * B, dtype=np.uint32)
    for b, lst in enumerate(kw):
        kwk[b, :len(lst)] = lst
        kwc[b] = len(lst)
    kwk = np.ascontiguousarray(kwk[:, :kl]) if kl else kwk[:, :0].copy()
    out = idx.alloc_fused(B, o)
    lat = []
    uncert = 0
    e2e_prof = os.environ.get("RAGERA_BENCH_E2E_PROF") == "1"   # diagnosis only: per-kernel device time inside the e2e calls
    if e2e_prof:
        idx.profile_enable(True)
        idx.profile_read()
    for i in range(total):
        lo = (i % n_pool) * B
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        r = idx.hybrid_raw(Qp[lo:lo + B], o, kwk[lo:lo + B], kwc[lo:lo + B], out)
        t1 = time.perf_counter()
        if i >= warmup:
            lat.append(t1 - t0)
            uncert += int(B - r.certified.sum())
    e2e_s = float(np.sum(lat))
    e2e_kernel_ms = None
    if e2e_prof:
        pr = idx.profile_read()
        idx.profile_enable(False)
        e2e_kernel_ms = {k: v[0] / total for k, v in pr.items() if v[1]}
    if world > 1:
        import torch

        t = torch.tensor([e2e_s], dtype=torch.float64, device=f"cuda:{device}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    cap = w["vector_top_k"] + w["keyword_limit"] + w.get("fresh_limit", 0)
    h2d = B * (d * 4 + w["keyword_limit"] * 8 + 4)
    d2h = B * (4 + 1 + 4 + 1 + cap * (8 + 8 + 1 + 1) + w["vector_top_k"] * 16)

    hbm_peak, tf_sus, tf_burst, peak_src = peaks()
    if tensor:
        k_ms, k_n = prof["tensor"]
        flops = 2.0 * n_local * idx_ld(d) * B                                 # dot products only
        ach = flops / (k_ms / max(k_n, 1) * 1e-3) / 1e12 if k_n else None
        op_bytes = 2 if (dt == N.BF16 or shadow) else 4                       # 16-bit operand, or fp32 rows read as tf32
        roof = {"bound": "tensor", "achieved": ach, "peak": tf_sus, "unit": "TFLOP/s", "frac": (ach / tf_sus) if ach else None,
                "traffic": None, "peak_source": peak_src + " bf16_tflops_sustained (kernel timed inside a long step)",
                "frac_of_burst_peak": (ach / tf_burst) if ach else None,
                "kernel": "k2_pair (tcgen05 cta_group::2 kind::f16 GEMM, fp16 queries x " + ("bf16 corpus rows" if dt == N.BF16 else
                          f"{shadow} shadow rows" if shadow else "tf32") + " + fused top-K' epilogue)", "algorithmic_flops_per_launch": flops,
                "algorithmic_bytes_per_launch": int(n_local * idx_ld(d) * op_bytes), "avg_launch_ms": k_ms / max(k_n, 1),
                "launches_timed": int(k_n)}
        if op_bytes == 4:
            # tf32 runs at half the bf16 tensor rate and streams 4 B/element: report the HBM side as well
            roof["operand"] = "fp32 corpus read as tf32 (kind::tf32, peak = half of the bf16 figure)"
            roof["peak"] = tf_sus / 2
            roof["frac"] = (ach / (tf_sus / 2)) if ach else None
            roof["hbm_achieved_GBps"] = (n_local * idx_ld(d) * 4) / (k_ms / max(k_n, 1) * 1e-3) / 1e9 if k_n else None
            roof["hbm_frac_of_measured"] = roof["hbm_achieved_GBps"] / hbm_peak if k_n else None
        tr = traffic_ratio("k2_pair", name)
        if tr:
            roof["traffic"] = int(tr * n_local * idx_ld(d) * 2)
            roof["traffic_source"] = "profiles/traffic.json (ncu --set full dram bytes / algorithmic, scaled to this launch)"
    else:
        k1_ms, k1_n = prof["stream"]
        bytes_per_launch = n_local * idx_ld(d) * (4 if (dt == N.F32 and not on_shadow) else 2) * B   # K1 streams the shard (or its fp16 shadow) once per query
        achieved = bytes_per_launch / (k1_ms / max(k1_n, 1) * 1e-3) / 1e9 if k1_n else None
        roof = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": (achieved / hbm_peak) if achieved else None, "traffic": None, "peak_source": peak_src + " hbm_gbs",
                "kernel": "k1_stream_f16n (fused GEMV over the fp16 shadow of the normalised rows + top-K')" if on_shadow else "k1_stream (fused cosine GEMV + top-K')",
                "algorithmic_bytes_per_launch": int(bytes_per_launch),
                "avg_launch_ms": k1_ms / max(k1_n, 1), "launches_timed": int(k1_n),
                "frac_of_nominal_8TBps": (achieved / 8000.0) if achieved else None}
        tr = traffic_ratio("k1_stream_f16n" if on_shadow else "k1_stream", name)
        if tr:
            roof["traffic"] = int(tr * bytes_per_launch)
            roof["traffic_source"] = "profiles/traffic.json (ncu --set full dram bytes / algorithmic, scaled to this launch)"
    res = {
        "value": steps * B * replicas / (ms * 1e-3), "ms_per_step": ms / steps, "gpu_launches": int(launches), "clocks": clocks,
        "e2e": {"value": steps * B * replicas / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "latency_ms_p50": float(np.median(lat) * 1e3), "latency_ms_p99": float(np.percentile(lat, 99) * 1e3),
                "timing": "host wall clock around the synchronous call (pinned host buffers; escalations included)",
                "uncertified_after_escalation": uncert},
        "roofline": roof,
        "kernel_ms_per_step": {k: v[0] / steps for k, v in prof.items() if v[1]},
        **({"e2e_kernel_ms_per_call": e2e_kernel_ms} if e2e_kernel_ms else {}),
        "certified": {"timed_steps": int(timed_certified), "timed_steps_of": int(timed_queries),
                      "note": "device-side count over ALL timed steps of `value` (first pass, no escalation there); the e2e "
                              "calls escalate uncertified queries inside the timed call",
                      "setup_queries": certified_setup, "of": n_pool * B, "last_step": int(last.certified.sum()), "last_step_of": B},
    }
    if timed_certified != timed_queries:
        res["certified"]["warning"] = (f"{timed_queries - timed_certified} of the {timed_queries} queries behind `value` were NOT certified by the first pass: "
                                       "the product call (e2e) re-runs them on a stronger path, the device-timed loop does not")
    if parity:
        res["parity_check"] = parity
    if first_pass:
        res["first_pass_certification"] = first_pass
    if shard_rows:
        res["shard_rows"] = shard_rows
    if world > 1:
        # every rank's own kernel times: the step ends when the SLOWEST rank's scoring kernel does (K5 waits for
        # every rank's records), so the spread between ranks is what the exchange wait in "fuse" is made of
        per_rank = [None] * world
        dist.all_gather_object(per_rank, {k: round(v, 4) for k, v in res["kernel_ms_per_step"].items()})
        res["per_rank_kernel_ms"] = per_rank
        res["exchange"] = "none: replicas (the whole corpus on every GPU, the batch split over the ranks)" if replicas > 1 else {"p2p": "peer-to-peer mailboxes (CUDA IPC over NVLink), written and awaited inside the fusion kernel; "
                                  "host-driven bootstrap, no NCCL call in the library",
                           "nccl": "ncclAllGather of the local top-k records before the fusion kernel"}[getattr(idx, "exchange", "p2p")]
    if do_cpu and rank == 0:
        # reference-faithful: ONE thread (the reference is single-threaded JS), full stable sort
        sample_rows = min(rows, 1_000_000)
        # about 10-15 s of single-thread CPU work: 4 queries x 1M rows (2.7 s each), or 200 queries of a 10k-row corpus
        n_q = 4 if sample_rows >= 500_000 else (40 if sample_rows >= 100_000 else 200)
        qps, per_query, sample, info = cpu_reference_leg(w, n_q, 0, 1, sample_rows)   # queries/s (batching does not help a scalar loop)
        res["cpu_baseline"] = {"value": qps, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                               "ms_per_query": per_query * 1e3, "extrapolated": info}
    if sharded:
        leave_exchange(dist, idx)
    idx.close()
    return res


def parity_check(w, gen, Qlast, last, d, bf16, n_queries=2, window=100_000):
    """Outside every timed region, on rank 0: the LAST timed step's answer for a few queries against the CPU oracle (the checker,
    never the thing measured) — (1) every reported cosine is the oracle's fp64 value for that chunk, bit for bit; (2) an oracle
    scan of a window of the corpus around the best hit finds no chunk that should have been in the answer and is not, and lists
    the returned chunks of that window in the returned order. Works for any shard count: ids are global."""
    import oracle

    g = oracle.GenDesc.from_buffer_copy(bytes(gen))
    odt = oracle.BF16 if bf16 else oracle.F32
    rows, k = w["rows"], w["vector_top_k"]
    out = {"queries": 0, "scores_bit_equal": True, "window_rows": min(window, rows), "window_consistent": True}
    for b in range(min(n_queries, len(Qlast))):
        r = last.row(b)
        ids, sc = [int(i) for i in r["vec_ids"]], [float(x) for x in r["vec_scores"]]
        if not ids:
            continue
        out["queries"] += 1
        for i, x in zip(ids, sc):
            if oracle.cosine(Qlast[b], oracle.gen_rows(g, i, 1, d, dtype=odt)[0]) != x:
                out["scores_bit_equal"] = False
        lo = max(0, min(rows - out["window_rows"], ids[0] - out["window_rows"] // 2))
        wi, ws = oracle.topk_generated(g, odt, lo, out["window_rows"], d, Qlast[b], k)
        inside = [i for i in ids if lo <= i < lo + out["window_rows"]]
        kth = (sc[-1], ids[-1])
        full = len(ids) == k
        must = [int(i) for i, x in zip(wi, ws) if x >= w["min_score"] and (not full or (x, -int(i)) > (kth[0], -kth[1]))]
        if [int(i) for i in wi[:len(inside)]] != inside or any(i not in ids for i in must):
            out["window_consistent"] = False
    out["ok"] = bool(out["queries"] > 0 and out["scores_bit_equal"] and out["window_consistent"])
    return out


def idx_ld(d):
    return (d + 255) & ~255


def run_ours(args):
    import rag_era_b200 as rb
    from rag_era_b200 import _native as N

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    # rank 0 prints exactly ONE JSON line on stdout: anything native code prints meanwhile (NCCL's version
    # banner, for one) is sent to stderr by pointing fd 1 at fd 2 until the line is ready
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N.load()
    w = WORKLOADS[args.workload]
    steps = args.steps or (200 if w["rows"] >= 1_000_000 else 1000)
    warmup = args.warmup if args.warmup is not None else 10
    warmup = max(3, warmup)
    res = measure_workload(rb, N, w, args.workload, steps, warmup, dist, rank, world, local, do_cpu=(world == 1))
    extra = {}
    if not args.no_extra:
        # the other BASELINE.json configs in the same line: configs[1] (c2, c2b), configs[0] (c1) and configs[3] (c4) on
        # one GPU; configs[4] (c5, 50M x 1536 bf16 batch 1024, row-sharded) at EVERY N so the scaling run carries its curve
        # (c1 is latency-bound — 45 us steps that follow the SM clock: 400 warm-up steps, 20 ms, let the clock come up after
        # the idle of the setup before the timed region starts)
        plan = [("c2", 200, 5), ("c2b", 30, 5), ("c1", 500, 400), ("c4", 20, 3)] if world == 1 else [("c2b", 30, 5), ("c2br", 30, 5)]
        plan.append(("c5", 10, 3))
        plan.append(("c3s", 200, 10))   # labelled extra mode: batch 1 on the fp16 shadow (never the headline)
        for name, e_steps, e_warm in plan:
            if name == args.workload:
                continue
            try:
                r = measure_workload(rb, N, WORKLOADS[name], name, e_steps, e_warm, dist, rank, world, local, do_cpu=False)
                extra[name] = {"workload": WORKLOADS[name]["desc"], "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": e_steps,
                               "warmup": e_warm, "ms_per_step": r["ms_per_step"], "e2e": r["e2e"], "roofline": r["roofline"],
                               "kernel_ms_per_step": r["kernel_ms_per_step"], "certified": r["certified"], "clocks": r["clocks"],
                               "gpu_launches": r["gpu_launches"]}
                for key in ("first_pass_certification", "per_rank_kernel_ms", "exchange", "shard_rows", "parity_check"):
                    if key in r:
                        extra[name][key] = r[key]
            except Exception as e:          # e.g. out of memory on a smaller part: say so instead of losing the line
                extra[name] = {"workload": WORKLOADS[name]["desc"], "error": f"{type(e).__name__}: {e}"[:300]}
    if rank == 0:
        line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config_of(w, args.workload, world), "clocks": res["clocks"],
                "e2e": res["e2e"], "gpu_launches": res["gpu_launches"], "roofline": res["roofline"],
                "kernel_ms_per_step": res["kernel_ms_per_step"], "certified": res["certified"],
                "arithmetic": "fp32 scoring selects K' candidates; fp64 reference-order rescoring decides ids/scores/ties"}
        if w.get("path") == "shadow_stream":
            line["dtype"] = "f16"   # selection streams fp16 rows against the fp32 query (fp32 accumulate); ids and scores are decided in fp64
        if res["roofline"].get("bound") == "tensor":
            # the batched path selects on tcgen05 products of 16-bit (fp16 queries x fp16/bf16 rows) or tf32 operands;
            # ids and scores are still decided in fp64
            line["dtype"] = "tf32" if "operand" in res["roofline"] else ("f16xbf16" if w["dtype"] == "bf16" or w.get("shadow") == "bf16" else "f16")
        for key in ("per_rank_kernel_ms", "exchange", "e2e_kernel_ms_per_call", "first_pass_certification", "shard_rows", "parity_check"):
            if key in res:
                line[key] = res[key]
        if "cpu_baseline" in res:
            line["cpu_baseline"] = res["cpu_baseline"]
        if extra:
            line["extra"] = extra
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-extra", action="store_true", help="skip the c2/c1 extra measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

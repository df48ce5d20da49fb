"""bench.py's contract, as far as it can be checked without a GPU: the reference arm prints ONE JSON line with the keys
the driver reads (and the tier's additions: impl, cpu_baseline, e2e with zero copy bytes), under torchrun only rank 0
prints; our own arm refuses to run without a device instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, timeout=300):
    return subprocess.run([sys.executable] + args, cwd=ROOT, capture_output=True, text=True, timeout=timeout)


def check_reference_line(stdout, n_gpus):
    lines = [l for l in stdout.splitlines() if l.strip()]
    assert len(lines) == 1, stdout                                   # exactly one line, from rank 0
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == n_gpus and d["steps"] == 1 and d["warmup"] == 1
    assert d["metric"] == "hybrid_search_qps" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] == d["value"] and "rows" in cb["sample"]   # variant A: one thread
    ex = d["extrapolated"]
    assert ex["threads"] == 1 and ex["sample_rows"] <= ex["rows"] and ex["scale_scoring"] >= 1 and ex["scale_sort_nlogn"] >= ex["scale_scoring"]
    assert d["all_threads_courtesy"]["cores"] >= 1 and d["all_threads_courtesy"]["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    return d


def test_reference_arm_line():
    r = run(["bench.py", "--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "1"])
    assert r.returncode == 0, r.stderr
    d = check_reference_line(r.stdout, 1)
    assert d["config"]["name"] == "c1" and d["config"]["rows"] == 10_000


def test_reference_arm_under_torchrun_prints_once():
    r = run(["-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
             "--master-port", str(29700 + os.getpid() % 200), "bench.py", "--impl", "reference", "--workload", "c1", "--gpus", "2",
             "--steps", "1", "--warmup", "1"])
    assert r.returncode == 0, r.stderr
    check_reference_line(r.stdout, 2)


def test_our_arm_refuses_without_a_device(native):
    if native.load().rag_device_count() > 0:
        pytest.skip("a GPU is visible")
    r = run(["bench.py", "--workload", "c1", "--no-extra", "--steps", "1", "--warmup", "1"])
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "no CPU fallback" in r.stderr

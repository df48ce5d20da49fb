"""The multi-rank path under the driver's ONE-GPU `pytest -m gpu`: two (and three) ranks as separate processes on
device 0, row shards, mailbox exchange over CUDA IPC, results equal to the oracle over the whole corpus; plus the
timeout / re-bootstrap behaviour of the exchange (tests/tools/ranks_one_gpu.py)."""
import os
import subprocess
import sys
import tempfile

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 3])
def test_ranks_sharing_one_gpu_equal_the_oracle(native, world):
    assert native.load().rag_device_count() > 0
    env = dict(os.environ)
    with tempfile.TemporaryDirectory(prefix="ragera_ranks_") as scratch:
        procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "tools", "ranks_one_gpu.py"), str(r), str(world), scratch],
                                  stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env) for r in range(world)]
        outs = []
        try:
            for p in procs:
                outs.append(p.communicate(timeout=900)[0])
        finally:
            for p in procs:
                if p.poll() is None:
                    p.kill()
        for r, (p, o) in enumerate(zip(procs, outs)):
            assert p.returncode == 0 and "parity OK" in o, f"rank {r} (exit {p.returncode}):\n{o[-3000:]}"

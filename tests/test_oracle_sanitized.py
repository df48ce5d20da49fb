"""The checker itself under ASAN + UBSAN: oracle/oracle.c rebuilt with sanitizers and the known-answer / property tests re-run
against that build in a subprocess (LD_PRELOAD of the sanitizer runtimes, since the host process is Python)."""
import os
import shutil
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_oracle_kats_against_a_sanitized_build(tmp_path):
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    so = str(tmp_path / "liboracle_asan.so")
    r = subprocess.run(["gcc", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-ffp-contract=off", "-pthread", "-fPIC",
                        "-shared", os.path.join(ROOT, "oracle", "oracle.c"), "-o", so, "-lm"], capture_output=True, text=True)
    if r.returncode != 0 and "sanitize" in r.stderr:
        pytest.skip("sanitizer runtimes are not installed")
    assert r.returncode == 0, r.stderr[-1500:]
    rt = [subprocess.run(["gcc", "-print-file-name=" + n], capture_output=True, text=True).stdout.strip() for n in ("libasan.so", "libubsan.so")]
    if not all(os.path.isabs(p) and os.path.exists(p) for p in rt):
        pytest.skip("sanitizer runtimes are not installed")
    script = tmp_path / "run.py"
    script.write_text(textwrap.dedent(f"""
        import sys
        sys.path.insert(0, {ROOT!r})
        import oracle
        oracle._LIB_PATH = {so!r}
        oracle.build = lambda force=False: oracle._LIB_PATH
        import pytest
        sys.exit(pytest.main(["-x", "-q", "-p", "no:cacheprovider", {os.path.join(ROOT, "tests", "test_oracle_kat.py")!r},
                              {os.path.join(ROOT, "tests", "test_oracle_properties.py")!r}]))
    """))
    env = dict(os.environ, LD_PRELOAD=" ".join(rt), ASAN_OPTIONS="detect_leaks=0")
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=900, env=env, cwd=str(tmp_path))
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    assert "AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr, r.stderr[-2000:]

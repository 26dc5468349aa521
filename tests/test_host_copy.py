"""The multi-threaded ingest copier of the streaming path (csrc/host_copy.h) is host-only code: built with the
system compiler and run here, no GPU (tests/cpp/host_copy_test.cpp; also under ThreadSanitizer when available)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_copy_test.cpp")
OUT = os.path.join(ROOT, "tests", "cpp", "build")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _run(extra, name):
    os.makedirs(OUT, exist_ok=True)
    exe = os.path.join(OUT, name)
    subprocess.check_call([CXX, "-std=c++14", "-O2", "-pthread"] + extra + [SRC, "-o", exe])
    return subprocess.run([exe], capture_output=True, text=True, timeout=600)


def test_parallel_copier_is_exact():
    res = _run([], "host_copy_test")
    assert res.returncode == 0 and "ALL HOST COPY TESTS PASSED" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]


def test_parallel_copier_under_thread_sanitizer():
    try:
        res = _run(["-fsanitize=thread", "-g"], "host_copy_test_tsan")
    except subprocess.CalledProcessError:
        pytest.skip("ThreadSanitizer runtime not available for this compiler")
    if "FATAL: ThreadSanitizer" in res.stderr or "unexpected memory mapping" in res.stderr:
        pytest.skip("ThreadSanitizer cannot run in this container")
    assert "WARNING: ThreadSanitizer" not in res.stderr, res.stderr[-4000:]
    assert res.returncode == 0 and "ALL HOST COPY TESTS PASSED" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]

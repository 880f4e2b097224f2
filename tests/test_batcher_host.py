"""Host logic of the micro-batching queue (oramacore_b200/csrc/batcher.h) with a fake executor:
tests/batcher_test.cpp is compiled with g++ (no CUDA) and run with 16 submitting threads; it
fails unless every caller receives exactly its own query's result and queries were coalesced."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("flags", [["-O2"], ["-O1", "-g", "-fsanitize=thread"]])
def test_batcher_merge_scatter_under_concurrency(tmp_path, flags):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    exe = str(tmp_path / "batcher_test")
    r = subprocess.run(["g++", "-std=c++17", *flags, "-pthread", "-I", ROOT, os.path.join(ROOT, "tests", "batcher_test.cpp"), "-o", exe],
                       capture_output=True, text=True)
    if r.returncode != 0 and "-fsanitize=thread" in flags:
        pytest.skip("ThreadSanitizer runtime not available: " + r.stderr[-200:])
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout[-500:], r.stderr[-2000:])
    assert "bad=0" in r.stdout

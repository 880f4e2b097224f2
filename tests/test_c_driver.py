"""A plain C99 program that includes only include/oramacore_b200.h (tests/c_driver/driver.c):
  * CPU: the header compiles as strict C99 and the program links against liboramacore_b200.so;
  * GPU: it runs one hybrid oc_search on the committed case tests/golden/c_driver_case.bin (generated from the
    oracle by tests/golden/make_c_driver_case.py) and every hit must equal the oracle's."""
import os
import shutil
import subprocess

import pytest

import oramacore_b200 as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_driver", "driver.c")


def _build(tmp_path):
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    ob.build()
    exe = str(tmp_path / "c_driver")
    libdir = os.path.dirname(ob.SO_PATH)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
                        "-L", libdir, "-l:liboramacore_b200.so", f"-Wl,-rpath,{libdir}", "-lm"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


def test_header_is_c99_and_the_driver_links(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)          # no argument: usage, before any device call
    assert r.returncode == 64 and "usage" in r.stderr


@pytest.mark.gpu
def test_c_driver_matches_the_committed_oracle_answers(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "c_driver_case.bin")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout[-500:], r.stderr[-2000:])
    assert "mismatches 0" in r.stdout

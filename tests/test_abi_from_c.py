"""The drop-in boundary is a C ABI: a C99 translation unit that includes include/amc_b200.h compiles without warnings
under -pedantic, links against the in-tree library and gets the same answers as the ctypes binding."""
import os
import shutil
import subprocess

import pytest

from vit_vs_raw_iq_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_header_is_c99_and_library_links_from_plain_c(tmp_path):
    exe = str(tmp_path / "abi_host")
    libdir = os.path.dirname(_lib.LIB_PATH)
    cc = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                         os.path.join(ROOT, "tests", "abi_host.c"), "-o", exe, "-L", libdir, "-lamc_b200",
                         "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr
    run = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, run.stdout + run.stderr
    lines = run.stdout.strip().splitlines()
    assert lines[0] == f"abi {_lib.ABI_VERSION}"
    assert lines[1] == "T 65 total 1985216"            # cfg-1: 1,985,163 parameters + alignment padding of the blob
    assert "must be divisible by n_head" in lines[2]   # R/training/train.py:132-133

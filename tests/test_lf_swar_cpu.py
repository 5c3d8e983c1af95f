"""CPU tier: the four-lines-per-register loop-filter arithmetic of the batch filter kernel
(vp8_b200/csrc/cuda/lf_swar.h, compiled for the host with emulated SIMD primitives) against the scalar
edge filter as the reference states it (src/filter.cc:7-67,119-149)."""
import os
import shutil
import subprocess

import pytest

import helpers


def test_swar_edges_equal_scalar_edges(tmp_path):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "lf_swar_test")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", "-I" + os.path.join(helpers.ROOT, "vp8_b200", "csrc"),
                           os.path.join(helpers.ROOT, "tests", "native", "lf_swar_test.cc"), "-o", exe])
    run = subprocess.run([exe, "24"], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout + run.stderr
    lines, filtered, bad = (int(x) for x in run.stdout.split()[1::2])
    assert bad == 0 and lines >= 24_000_000 and filtered > lines // 5

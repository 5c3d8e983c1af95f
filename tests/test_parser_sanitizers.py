"""CPU tier: the host parser under AddressSanitizer + UBSan on corrupted streams (it reads untrusted
input: a compressed frame must never make it read or write out of bounds, whatever the parse mode)."""
import os
import shutil
import subprocess

import pytest

import helpers

SRC = os.path.join(helpers.ROOT, "tests", "native", "parser_fuzz.cc")
PARSER = os.path.join(helpers.ROOT, "vp8_b200", "csrc", "host", "frame_parser.cc")


def test_parser_fuzz_under_sanitizers(built, tmp_path):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "parser_fuzz")
    cmd = ["g++", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer",
           "-std=c++17", "-I" + os.path.join(helpers.ROOT, "include"), "-I" + os.path.join(helpers.ROOT, "vp8_b200", "csrc"),
           SRC, PARSER, "-o", exe]
    build = subprocess.run(cmd, capture_output=True, text=True)
    if build.returncode != 0:
        pytest.skip("sanitizer build unavailable: " + build.stderr[-300:])
    for k, args in enumerate(("--width 96 --height 80 --frames 6 --seed 5 --log2-parts 2 --pct-split 30 --pct-intra 20",
                              "--width 176 --height 144 --frames 5 --seed 11 --log2-parts 3 --pct-split 20 --pct-intra 30 --key-interval 2")):
        ivf = str(tmp_path / f"s{k}.ivf")
        open(ivf, "wb").write(helpers.synth_stream(args))
        run = subprocess.run([exe, ivf, "60"], capture_output=True, text=True, timeout=600)
        assert run.returncode == 0, run.stderr[-2000:]
        assert "ERROR" not in run.stderr and "runtime error" not in run.stderr, run.stderr[-2000:]
        ok, rejected = (int(x) for x in run.stdout.split()[1::2])
        assert ok > 100 and rejected > 50

#!/usr/bin/env python3
"""Kernel-only replay of the bench workload without torch (short start-up: for ncu captures and A/B runs).

  prof_replay.py --streams 64 --frames 8 --passes 3 [--size 1920x1080]

Parses S synthetic streams (bench.py's generator settings) on the host, uploads the parsed frames, replays
them `passes` times and prints the per-kernel-class device times of the last pass as one JSON line."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vp8_b200

SYNTH = os.path.join(ROOT, "vp8_b200", "_lib", "vp8synth")
ARGS = "--log2-parts 2 --q 40 --lf 24 --pct-skip 55 --coef-density 3 --pct-empty-block 80"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=64)
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--passes", type=int, default=3)
    ap.add_argument("--size", default="1920x1080")
    ap.add_argument("--synth-args", default=ARGS)
    ap.add_argument("--engines", type=int, default=1, help="split the streams over this many engines (own CUDA streams), "
                    "time steps submitted round-robin: kernels of different engines overlap")
    ap.add_argument("--serial-setup", action="store_true", help="generate the streams one after the other (under ncu the "
                    "thread pool that runs the generator takes the process down)")
    ap.add_argument("--parse", default="host", choices=["host", "tokens", "device"],
                    help="what the replay includes: host = reconstruction kernels only; tokens = + the DCT token kernel; "
                         "device = + the macroblock-header kernel (everything behind the frame headers on the GPU)")
    a = ap.parse_args()
    w, h = (int(x) for x in a.size.split("x"))
    tmp = tempfile.mkdtemp()

    def make(k):
        p = os.path.join(tmp, f"s{k}.ivf")
        subprocess.check_call([SYNTH, "--width", str(w), "--height", str(h), "--frames", str(a.frames), "--seed", str(7122 + k),
                               "--out", p] + a.synth_args.split())
        return vp8_b200.read_ivf(p)[1]

    if a.serial_setup:
        payloads = [make(k) for k in range(a.streams)]
    else:
        with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
            payloads = list(ex.map(make, range(a.streams)))
    if a.engines > 1:
        return multi_engine(a, payloads)
    eng = vp8_b200.Engine(0)
    eng.set_timing(True)
    dec = vp8_b200.BatchDecoder(eng, a.streams, pinned=False, tokens_on_device=a.parse == "tokens", device_parse=a.parse == "device")
    resident = []
    for t in range(a.frames):
        frames = dec.parse_into([vp8_b200.ParsedFrame(pinned=False) for _ in range(a.streams)], [p[t] for p in payloads],
                                list(range(a.streams)))
        for f in frames:
            eng.upload(f, release_host=True)
        resident.append(frames)
    out = None
    import time
    for _ in range(a.passes):
        eng.timers(reset=True)
        eng.set_timing(False)
        eng.sync()
        t0 = time.perf_counter()
        for t in range(a.frames):
            eng.reconstruct_batch(dec.streams, resident[t])
        eng.sync()
        wall = time.perf_counter() - t0
        eng.set_timing(True)
        for t in range(a.frames):
            eng.reconstruct_batch(dec.streams, resident[t])
        tm = eng.timers(reset=True)
        out = {"streams": a.streams, "frames": a.frames, "size": a.size, "ms_inter": tm.ms_inter, "ms_intra": tm.ms_intra,
               "ms_filter": tm.ms_filter, "ms_border": tm.ms_border, "ms_parse_kernels": tm.ms_tokens, "parse": a.parse, "filter_ms_per_launch": tm.ms_filter / max(1, tm.launches_filter),
               "inter_ms_per_launch": tm.ms_inter / max(1, tm.launches_inter),
               "frames_per_s": tm.frames / max(1e-9, (tm.ms_inter + tm.ms_intra + tm.ms_filter + tm.ms_border) / 1e3),
               "filter_mode": os.environ.get("VP8R_FILTER", "auto"), "wall_ms_untimed_pass": wall * 1e3,
               "wall_frames_per_s": a.streams * a.frames / wall}
    sums = eng.checksum_batch(dec.streams)
    out["checksum_xor"] = 0
    for s in sums:
        out["checksum_xor"] ^= s
    print(json.dumps(out))
    for fr in resident:
        for f in fr:
            f.close()
    dec.close()
    eng.close()


def multi_engine(a, payloads):
    import time
    per = a.streams // a.engines
    engs, decs, res = [], [], []
    for e in range(a.engines):
        eng = vp8_b200.Engine(0)
        dec = vp8_b200.BatchDecoder(eng, per, pinned=False)
        mine = payloads[e * per:(e + 1) * per]
        resident = []
        for t in range(a.frames):
            frames = dec.parse_into([vp8_b200.ParsedFrame(pinned=False) for _ in range(per)], [p[t] for p in mine], list(range(per)))
            for f in frames:
                eng.upload(f, release_host=True)
            resident.append(frames)
        engs.append(eng)
        decs.append(dec)
        res.append(resident)
    for _ in range(a.passes):
        for eng in engs:
            eng.sync()
        t0 = time.perf_counter()
        for t in range(a.frames):
            for e in range(a.engines):
                engs[e].reconstruct_batch(decs[e].streams, res[e][t])
        for eng in engs:
            eng.sync()
        wall = time.perf_counter() - t0
    x = 0
    for e in range(a.engines):
        for s in engs[e].checksum_batch(decs[e].streams):
            x ^= s
    print(json.dumps({"streams": a.streams, "engines": a.engines, "frames": a.frames, "wall_ms_untimed_pass": wall * 1e3,
                      "wall_frames_per_s": per * a.engines * a.frames / wall, "filter_mode": os.environ.get("VP8R_FILTER", "auto"),
                      "checksum_xor": x}))


if __name__ == "__main__":
    main()

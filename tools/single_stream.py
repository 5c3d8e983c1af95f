#!/usr/bin/env python3
"""BASELINE config 3 / 5 as a latency case: ONE synthetic stream decoded frame by frame (each frame waits
for the previous one), host parse vs device-side parse.  Prints frames/s and the device time per kernel
class per frame."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, vp8_b200
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import helpers

eng = vp8_b200.Engine(0)
eng.set_timing(True)
for name, args in (("1080p", bench.SYNTH_ARGS), ("2160p", bench.SYNTH_ARGS.replace("1920", "3840").replace("1080", "2160"))):
    _, payloads = vp8_b200.read_ivf(helpers.synth_stream(args + " --seed 7122"))
    for mode in ("host", "device"):
        for rep in range(2):  # first repetition warms up allocations
            ps, st = vp8_b200.Parser(), eng.open_stream()
            if mode == "device":
                ps.set_defer_modes(True)
            fr = vp8_b200.ParsedFrame(pinned=True)
            eng.timers(reset=True)
            t0 = time.perf_counter()
            for p in payloads:
                ps.parse(p, out=fr)
                eng.reconstruct_batch([st], [fr])
                eng.sync()  # the next frame's parse may reuse `fr`; the output is complete here
            dt = time.perf_counter() - t0
            tm = eng.timers(reset=True)
            st.close()
            fr.close()
        n = len(payloads)
        print(f"{name} {mode:6s}: {n / dt:7.1f} frames/s  ({1e3 * dt / n:6.2f} ms/frame; device ms/frame: "
              f"parse {tm.ms_tokens / n:5.2f} inter {tm.ms_inter / n:5.2f} intra {tm.ms_intra / n:5.2f} filter {tm.ms_filter / n:5.2f})",
              flush=True)
eng.close()

#!/usr/bin/env python3
"""Development probe: where the end-to-end time goes (parse mode x read-back on/off)."""
import os, sys, time, tempfile, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, vp8_b200
from concurrent.futures import ThreadPoolExecutor

S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
DISTINCT = min(S, int(os.environ.get("PROBE_DISTINCT", "512")))  # further streams repeat the first ones
tmp = tempfile.mkdtemp()
paths = [os.path.join(tmp, f"s{k}.ivf") for k in range(DISTINCT)]
with ThreadPoolExecutor(max_workers=os.cpu_count()) as ex:
    list(ex.map(lambda k: bench.synth_stream(7122 + k, paths[k]), range(DISTINCT)))
payloads = [vp8_b200.read_ivf(p)[1] for p in paths]
payloads = [payloads[k % DISTINCT] for k in range(S)]
shutil.rmtree(tmp, ignore_errors=True)
stream = torch.cuda.Stream(priority=int(os.environ.get('PROBE_PRIO', '0')))  # -1: high
eng = vp8_b200.Engine(0, cuda_stream=stream.cuda_stream)
eng.set_timing(True)
DEPTH = int(sys.argv[4]) if len(sys.argv) > 4 else 4
ring = [torch.empty((S, bench.FRAME_BYTES), dtype=torch.uint8, pin_memory=True) for _ in range(DEPTH)]
packed = (tuple(r.data_ptr() for r in ring), bench.FRAME_BYTES)
MODES = sys.argv[2].split(",") if len(sys.argv) > 2 else ["host", "tokens", "device"]
RB = [bool(int(x)) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [False, True]
for mode in MODES:
    for rb in RB:
        dec = vp8_b200.BatchDecoder(eng, S, parse_threads=os.cpu_count(), pinned=True,
                                    tokens_on_device=mode == "tokens", depth=DEPTH,
                                    device_parse=(True if mode == "device" else float(mode[3:]) if mode.startswith("mix") else False))
        dec.decode(payloads, out_packed=packed if rb else None)
        torch.cuda.synchronize()
        eng.timers(reset=True)
        PASSES = int(os.environ.get("PROBE_PASSES", "1"))  # > 1: passes follow each other without draining
        t0 = time.perf_counter()
        for k in range(PASSES):
            dec.reset()
            dec.decode(payloads, out_packed=packed if rb else None, drain=k + 1 == PASSES)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / PASSES
        tm = eng.timers(reset=True)
        print(f"{mode:7s} readback={rb!s:5s} {S*30/dt:8.0f} fps  wall {dt*1e3:7.0f} ms  tokens {tm.ms_tokens:6.0f} recon {tm.ms_inter+tm.ms_intra+tm.ms_filter+tm.ms_border:6.0f} h2d {tm.ms_h2d:5.0f} pack {tm.ms_d2h:5.0f}", flush=True)
        print("        host seconds:", {k: round(v, 3) for k, v in dec.host_seconds.items()}, flush=True)
        dec.close()
eng.close()

#!/usr/bin/env python3
"""Debug aid: decodes synthetic streams with the scalar and the batch (SWAR) loop filter and reports where
the frames differ.  usage: filter_debug.py [vp8synth args ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import helpers
import vp8_b200


def decode(data, mode):
    os.environ["VP8R_FILTER"] = mode
    eng = vp8_b200.Engine(0)
    _, payloads = vp8_b200.read_ivf(data)
    st = eng.open_stream()
    out = []
    for p in payloads:
        st.decode(p)
        out.append((st.read_frame(), st.dims()))
    st.close()
    eng.close()
    return out


def main():
    cases = sys.argv[1:] or ["--width 1920 --height 1080 --frames 3 --seed 7122 --log2-parts 2",
                             "--width 176 --height 144 --frames 24 --seed 7122",
                             "--width 640 --height 360 --frames 10 --seed 7 --segmentation 0 --lf-deltas 0 --pct-split 40",
                             "--width 1920 --height 96 --frames 8 --seed 8 --log2-parts 1"]
    for args in cases:
        data = helpers.synth_stream(args)
        a, b = decode(data, "scalar"), decode(data, "swar")
        print("==", args)
        for k, ((fa, (w, h)), (fb, _)) in enumerate(zip(a, b)):
            if fa == fb:
                print(f" frame {k}: equal")
                continue
            ya = np.frombuffer(fa, np.uint8)
            yb = np.frombuffer(fb, np.uint8)
            cw, ch = (w + 1) // 2, (h + 1) // 2
            for name, off, pw, ph in (("Y", 0, w, h), ("U", w * h, cw, ch), ("V", w * h + cw * ch, cw, ch)):
                pa = ya[off:off + pw * ph].reshape(ph, pw).astype(int)
                pb = yb[off:off + pw * ph].reshape(ph, pw).astype(int)
                d = np.argwhere(pa != pb)
                if len(d) == 0:
                    continue
                ys, xs = d[:, 0], d[:, 1]
                print(f" frame {k} plane {name}: {len(d)} pixels differ, rows {ys.min()}..{ys.max()} cols {xs.min()}..{xs.max()}")
                mb = pw // (16 if name == "Y" else 8) + 1
                n = 16 if name == "Y" else 8
                mbs = sorted({(int(y) // n, int(x) // n) for y, x in d})
                print("   macroblocks (r,c):", mbs[:24], "..." if len(mbs) > 24 else "")
                print("   first:", [(int(y), int(x), int(pa[y, x]), int(pb[y, x])) for y, x in d[:12]])
                print("   y%n histogram:", np.bincount(ys % n, minlength=n).tolist(), " x%n:", np.bincount(xs % n, minlength=n).tolist())


if __name__ == "__main__":
    main()

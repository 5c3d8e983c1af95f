"""Test-side helpers: golden MD5 lists, the oracle binding (oracle/ is test infrastructure and is
only ever loaded from here, from __graft_entry__.smoke() and from bench.py's CPU legs)."""
import ctypes as C
import glob
import hashlib
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VEC_DIR = os.path.join(ROOT, "tests", "golden", "vp8-test-vectors")
SYN_DIR = os.path.join(ROOT, "tests", "golden", "synthetic")
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
LIB_SO = os.path.join(ROOT, "vp8_b200", "_lib", "libvp8r.so")
REF_DECODE = os.path.join(ROOT, "oracle", "_ref", "decode")


def ensure_built():
    if not (os.path.exists(LIB_SO) and os.path.exists(ORACLE_SO)):
        subprocess.check_call([os.path.join(ROOT, "build.sh")], cwd=ROOT)


def vectors():
    """The 43 shipped test vectors, plus every *.ivf with a *.ivf.md5 beside it under $VP8_TEST_VECTORS (the 18
    vp80-00-comprehensive-* vectors are git-ignored in the reference, test/test_comprehensive.py:13-20: supply
    them there and every vector test picks them up)."""
    found = sorted(glob.glob(os.path.join(VEC_DIR, "*.ivf")))
    extra = os.environ.get("VP8_TEST_VECTORS")
    if extra and os.path.isdir(extra):
        have = {os.path.basename(p) for p in found}
        for p in sorted(glob.glob(os.path.join(extra, "**", "*.ivf"), recursive=True)):
            if os.path.exists(p + ".md5") and os.path.basename(p) not in have:
                found.append(p)
    return found


def golden_md5(ivf_path):
    """[(md5, width, height)] per shown frame; the size comes from the md5 line's file name
    (streams 1425 and 1436 change size mid-stream)."""
    out = []
    for line in open(ivf_path + ".md5"):
        md5, name = line.split()
        m = re.search(r"-(\d+)x(\d+)-(\d+)\.i420$", name)
        out.append((md5, int(m.group(1)), int(m.group(2))))
    return out


def i420_bytes(w, h):
    return w * h + 2 * ((w + 1) // 2) * ((h + 1) // 2)


def md5(b):
    return hashlib.md5(b).hexdigest()


class Oracle:
    """ctypes view of oracle/_build/liboracle.so (CPU restatement of the reconstruction path)."""

    def __init__(self):
        ensure_built()
        self.lib = C.CDLL(ORACLE_SO)
        self.lib.oracle_create.restype = C.c_void_p
        self.lib.oracle_destroy.argtypes = [C.c_void_p]
        self.lib.oracle_decode_frame.argtypes = [C.c_void_p, C.c_void_p]
        self.lib.oracle_decode_frame.restype = C.c_int
        self.lib.oracle_frame_bytes.argtypes = [C.c_void_p]
        self.lib.oracle_frame_bytes.restype = C.c_size_t
        self.lib.oracle_write_i420.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        self.lib.oracle_write_i420.restype = C.c_size_t
        self.lib.oracle_idct4x4.argtypes = [C.POINTER(C.c_int16)]
        self.lib.oracle_iwht4x4.argtypes = [C.POINTER(C.c_int16)]
        self.h = self.lib.oracle_create()

    def decode(self, parsed_frame):
        """parsed_frame: vp8_b200.ParsedFrame.  Returns the I420 bytes of the reconstructed frame."""
        d = parsed_frame.desc()
        rc = self.lib.oracle_decode_frame(self.h, C.byref(d))
        if rc != 0:
            raise RuntimeError(f"oracle_decode_frame: {rc}")
        n = self.lib.oracle_frame_bytes(self.h)
        buf = (C.c_uint8 * n)()
        self.lib.oracle_write_i420(self.h, buf, n)
        return bytes(buf)

    def close(self):
        if self.h:
            self.lib.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()


def oracle_decode_ivf(path_or_bytes):
    """Shown frames of a stream through host parser -> oracle."""
    import vp8_b200
    _, payloads = vp8_b200.read_ivf(path_or_bytes)
    ps, orc = vp8_b200.Parser(), Oracle()
    out = []
    for p in payloads:
        fr = ps.parse(p)
        img = orc.decode(fr)
        if fr.desc().hdr.show_frame:
            out.append(img)
        fr.close()
    orc.close()
    return out


def ref_decode_ivf(path):
    """The compiled, unmodified reference decoder (oracle/_ref/decode) on a file; raw YUV bytes."""
    import tempfile
    with tempfile.NamedTemporaryFile(suffix=".yuv") as t:
        subprocess.check_call([REF_DECODE, path, t.name])
        return open(t.name, "rb").read()


SYNTH_BIN = os.path.join(ROOT, "vp8_b200", "_lib", "vp8synth")


def synth_manifest():
    import json
    return json.load(open(os.path.join(SYN_DIR, "manifest.json")))


def synth_stream(args):
    """Runs vp8synth with `args` (str) and returns the IVF bytes."""
    import tempfile
    ensure_built()
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "s.ivf")
        subprocess.check_call([SYNTH_BIN] + args.split() + ["--out", p])
        return open(p, "rb").read()


def synth_golden(name):
    return [l.split()[0] for l in open(os.path.join(SYN_DIR, name + ".md5"))]

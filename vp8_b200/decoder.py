"""Host-side mirror of the reference's decode loop (src/decode.cc:50-77) over the C ABI.

    Parser        <-> BitstreamParser + ParserContext       (host only)
    Engine/Stream <-> DecodeFrame + RefreshRefFrames + YUV::WriteFrame on the GPU

There is no CPU implementation of the pixel path here: Engine() raises when libvp8r.so is missing
or no sm_100 GPU is visible.
"""
import ctypes as C

from . import _capi
from ._capi import check


class ParsedFrame:
    def __init__(self, pinned=False):
        self._lib = _capi.load()
        self.handle = self._lib.vp8r_frame_create(1 if pinned else 0)
        if not self.handle:
            raise MemoryError("vp8r_frame_create")

    def desc(self):
        d = _capi.FrameDesc()
        check(self._lib.vp8r_frame_get_desc(self.handle, C.byref(d)))
        return d

    def write_bitstream(self):
        """The frame as a VP8 key frame (the inverse of Parser.parse for what the encoder produces)."""
        size = C.c_size_t(0)
        rc = self._lib.vp8r_frame_write_bitstream(self.handle, None, 0, C.byref(size))
        if size.value == 0:
            check(rc)
        buf = (C.c_uint8 * size.value)()
        check(self._lib.vp8r_frame_write_bitstream(self.handle, buf, size.value, C.byref(size)))
        return bytes(buf[:size.value])

    def close(self):
        if self.handle:
            self._lib.vp8r_frame_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Parser:
    """One per stream: carries probability / segmentation state between frames."""

    def __init__(self):
        self._lib = _capi.load()
        self.handle = self._lib.vp8r_parser_create()

    def parse(self, payload, out=None, pinned=False):
        out = out or ParsedFrame(pinned=pinned)
        buf = (C.c_uint8 * len(payload)).from_buffer_copy(payload) if len(payload) else (C.c_uint8 * 1)()
        check(self._lib.vp8r_parser_parse(self.handle, buf, len(payload), out.handle))
        return out

    def reset(self):
        self._lib.vp8r_parser_reset(self.handle)

    def set_defer_tokens(self, on=True):
        """First partition only on the host; the DCT partitions travel with the frame and are
        decoded by the engine's token kernel (vp8r_frame_hdr.tokens_deferred)."""
        self._lib.vp8r_parser_set_defer_tokens(self.handle, 1 if on else 0)

    def set_defer_modes(self, on=True):
        """Frame headers only on the host; the per-macroblock syntax of the first partition and the
        DCT partitions are decoded by the engine's parse kernel (vp8r_frame_hdr.modes_deferred)."""
        self._lib.vp8r_parser_set_defer_modes(self.handle, 1 if on else 0)

    def close(self):
        if self.handle:
            self._lib.vp8r_parser_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    def __init__(self, device=0, cuda_stream=None):
        self._lib = _capi.load()
        h = C.c_void_p()
        check(self._lib.vp8r_engine_create(device, C.c_void_p(cuda_stream) if cuda_stream else None, C.byref(h)))
        self.handle = h
        self.device = device

    def open_stream(self):
        return Stream(self)

    def reconstruct_batch(self, streams, frames):
        n = len(streams)
        sa = (C.c_void_p * n)(*[s.handle for s in streams])
        fa = (C.c_void_p * n)(*[f.handle for f in frames])
        check(self._lib.vp8r_reconstruct_batch(self.handle, n, sa, fa))

    def encode_key_frames(self, streams, images, width, height, q_index, loop_filter_level=0, sharpness=0, bpred=True):
        """Row f4: one key frame per stream from cropped I420 images (bytes), encoded in a closed loop on the device.
        Returns the ParsedFrame objects (macroblock records + coefficient blocks; .write_bitstream() serialises them);
        every stream then holds the reconstructed frame a decoder would produce.  bpred: B_PRED (sixteen 4x4 modes per
        macroblock) competes with the four 16x16 modes (VP8R_ENC_BPRED)."""
        n = len(streams)
        need = width * height + 2 * ((width + 1) // 2) * ((height + 1) // 2)
        bufs = []
        for img in images:
            if len(img) != need:
                raise ValueError("image is not a cropped I420 frame of the given size")
            bufs.append((C.c_uint8 * need).from_buffer_copy(img))
        frames = [ParsedFrame(pinned=False) for _ in range(n)]
        sa = (C.c_void_p * n)(*[s.handle for s in streams])
        ia = (C.c_void_p * n)(*[C.cast(b, C.c_void_p) for b in bufs])
        fa = (C.c_void_p * n)(*[f.handle for f in frames])
        check(self._lib.vp8r_encode_key_frames(self.handle, n, sa, ia, width, height, q_index, loop_filter_level, sharpness, 1 if bpred else 0, fa))
        return frames

    def upload(self, frame, release_host=False):
        """Copies the frame's arrays to HBM; release_host frees the host copy afterwards."""
        check(self._lib.vp8r_frame_upload(self.handle, frame.handle))
        if release_host:
            check(self._lib.vp8r_frame_release_host(frame.handle))

    def checksum_batch(self, streams):
        n = len(streams)
        sa = (C.c_void_p * n)(*[s.handle for s in streams])
        out = (C.c_uint64 * n)()
        check(self._lib.vp8r_checksum_batch(self.handle, n, sa, out))
        return list(out)

    def read_batch(self, streams, ptrs, caps, async_=False):
        n = len(streams)
        sa = (C.c_void_p * n)(*[s.handle for s in streams])
        pa = (C.c_void_p * n)(*ptrs)
        ca = (C.c_size_t * n)(*caps)
        check(self._lib.vp8r_read_batch(self.handle, n, sa, pa, ca, 1 if async_ else 0))

    def read_batch_packed(self, streams, dst_ptr, stride, async_=False, layout="i420"):
        """Device-side crop + pack of every stream's latest frame, one contiguous D2H copy:
        frame i at dst_ptr + i*stride.  layout: "i420" (Y, U, V planes) or "nv12" (Y, interleaved UV)."""
        n = len(streams)
        sa = (C.c_void_p * n)(*[s.handle for s in streams])
        check(self._lib.vp8r_read_batch_packed_as(self.handle, n, sa, C.c_void_p(dst_ptr), stride, 1 if async_ else 0,
                                                  {"i420": 0, "nv12": 1}[layout]))

    def sync(self):
        check(self._lib.vp8r_engine_sync(self.handle))

    def set_timing(self, on):
        check(self._lib.vp8r_engine_set_timing(self.handle, 1 if on else 0))

    def timers(self, reset=False):
        t = _capi.Timers()
        check(self._lib.vp8r_engine_get_timers(self.handle, C.byref(t), 1 if reset else 0))
        return t

    def close(self):
        if self.handle:
            self._lib.vp8r_engine_destroy(self.handle)
            self.handle = None


class Stream:
    """One decoder instance: parser state + last/golden/altref surfaces resident in HBM."""

    def __init__(self, engine):
        self.engine = engine
        self._lib = engine._lib
        h = C.c_void_p()
        check(self._lib.vp8r_stream_open(engine.handle, C.byref(h)))
        self.handle = h

    def decode(self, payload):
        """One iteration of src/decode.cc:50-77.  Returns show_frame."""
        shown = C.c_int(0)
        buf = (C.c_uint8 * max(len(payload), 1)).from_buffer_copy(payload or b"\0")
        check(self._lib.vp8r_stream_decode(self.handle, buf, len(payload), C.byref(shown)))
        return bool(shown.value)

    def frame_bytes(self):
        return self._lib.vp8r_stream_frame_bytes(self.handle)

    def dims(self):
        w, h = C.c_int(), C.c_int()
        check(self._lib.vp8r_stream_dims(self.handle, C.byref(w), C.byref(h)))
        return w.value, h.value

    def read_frame(self):
        n = self.frame_bytes()
        buf = (C.c_uint8 * n)()
        check(self._lib.vp8r_stream_read_frame(self.handle, buf, n))
        return bytes(buf)

    def checksum(self):
        out = C.c_uint64()
        check(self._lib.vp8r_stream_checksum(self.handle, C.byref(out)))
        return out.value

    def close(self):
        if self.handle:
            self._lib.vp8r_stream_close(self.handle)
            self.handle = None


def decode_ivf(path_or_bytes, engine=None):
    """`./decode in.ivf out.yuv` (src/decode.cc:13-79): returns the shown frames as I420 bytes."""
    from .ivf import read_ivf
    own = engine is None
    engine = engine or Engine()
    _, payloads = read_ivf(path_or_bytes)
    st = engine.open_stream()
    out = []
    try:
        for p in payloads:
            if st.decode(p):
                out.append(st.read_frame())
    finally:
        st.close()
        if own:
            engine.close()
    return out

"""Multi-GPU partitioning.  The path shards with no exchange step: streams are independent, and a
single stream can be cut at key frames because a key frame resets the parser context
(src/bitstream_parser.cc:43-56) and refreshes all three reference buffers (:88-92,
src/loop.h:34-45) -- the same property src/display.cc:53-75 uses to seek.  No collective touches
pixel data; ranks only agree on who decodes what."""
from . import _capi


def stream_owner(stream_index, world_size):
    """BASELINE config 4: stream k -> GPU k mod G."""
    return stream_index % world_size


def streams_of_rank(n_streams, rank, world_size):
    return [k for k in range(n_streams) if stream_owner(k, world_size) == rank]


def split_at_key_frames(payloads):
    """BASELINE config 5: key-frame-delimited segments of one stream.  Returns [(first_frame_index,
    [payloads...])]; every segment starts with a key frame and decodes independently."""
    lib = _capi.load()
    segs = []
    for i, p in enumerate(payloads):
        if lib.vp8r_is_key_frame(p, len(p)) or not segs:
            segs.append((i, []))
        segs[-1][1].append(p)
    if segs and not lib.vp8r_is_key_frame(segs[0][1][0], len(segs[0][1][0])):
        raise ValueError("stream does not start with a key frame")
    return segs


def segments_of_rank(segments, rank, world_size):
    """Segment g -> GPU g mod G."""
    return [(g, s) for g, s in enumerate(segments) if g % world_size == rank]

"""GPU tier: the CUDA path through the C ABI against the golden MD5s and, byte for byte,
against the oracle on the same parsed frames."""
import os

import pytest

import helpers

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine(built):
    import vp8_b200
    e = vp8_b200.Engine(0)
    yield e
    e.close()


def first_diff(a, b):
    n = min(len(a), len(b))
    for i in range(n):
        if a[i] != b[i]:
            return i
    return n if len(a) != len(b) else -1


@pytest.mark.parametrize("ivf", helpers.vectors(), ids=lambda p: os.path.basename(p)[:-4])
def test_stream_decode_matches_golden_and_oracle(engine, ivf):
    """One stream at a time through vp8r_stream_decode / vp8r_stream_read_frame (src/decode.cc loop)."""
    import vp8_b200
    gold = helpers.golden_md5(ivf)
    _, payloads = vp8_b200.read_ivf(ivf)
    ps, orc = vp8_b200.Parser(), helpers.Oracle()
    st = engine.open_stream()
    shown = 0
    try:
        for k, p in enumerate(payloads):
            fr = ps.parse(p)
            want = orc.decode(fr)
            is_shown = st.decode(p)
            got = st.read_frame()
            assert is_shown == bool(fr.desc().hdr.show_frame)
            d = first_diff(got, want)
            assert d < 0, f"frame {k}: first differing byte {d} (of {len(want)})"
            assert st.checksum() == engine._lib.vp8r_checksum_i420(got, *st.dims())
            if is_shown:
                md5, w, h = gold[shown]
                assert (w, h) == st.dims()
                assert helpers.md5(got) == md5, f"shown frame {shown}"
                shown += 1
            fr.close()
        assert shown == len(gold)
    finally:
        st.close()
        orc.close()


def test_all_vectors_batched(engine):
    """BASELINE config 2: the whole suite as 43 concurrent streams, one batched launch per time step."""
    import vp8_b200
    vecs = helpers.vectors()
    payloads = [vp8_b200.read_ivf(v)[1] for v in vecs]
    golds = [helpers.golden_md5(v) for v in vecs]
    parsers = [vp8_b200.Parser() for _ in vecs]
    streams = [engine.open_stream() for _ in vecs]
    shown = [0] * len(vecs)
    try:
        for t in range(max(len(p) for p in payloads)):
            live = [i for i in range(len(vecs)) if t < len(payloads[i])]
            frames = [parsers[i].parse(payloads[i][t]) for i in live]
            engine.reconstruct_batch([streams[i] for i in live], frames)
            engine.sync()
            for i, fr in zip(live, frames):
                if fr.desc().hdr.show_frame:
                    md5, w, h = golds[i][shown[i]]
                    assert helpers.md5(streams[i].read_frame()) == md5, f"{os.path.basename(vecs[i])} frame {shown[i]}"
                    shown[i] += 1
                fr.close()
        assert shown == [len(g) for g in golds]
    finally:
        for s in streams:
            s.close()


@pytest.mark.parametrize("name", sorted(helpers.synth_manifest().keys()))
def test_synthetic_streams_match_reference_md5(engine, name):
    """Synthetic streams (bilinear, full-pixel, simple filter, hidden alt-ref, multi-partition,
    1080p) against MD5s produced by the unmodified reference decoder."""
    import vp8_b200
    m = helpers.synth_manifest()[name]
    frames = vp8_b200.decode_ivf(helpers.synth_stream(m["args"]), engine=engine)
    gold = helpers.synth_golden(name)
    assert len(frames) == len(gold)
    for k, (img, md5) in enumerate(zip(frames, gold)):
        assert helpers.md5(img) == md5, f"{name} shown frame {k}"

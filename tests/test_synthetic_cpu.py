"""CPU tier: synthetic streams.  Goldens = MD5s of the UNMODIFIED reference decoder's output
(tests/golden/make_synthetic_goldens.py); they pin the bilinear / full-pixel / simple-filter /
golden-copy paths that no shipped vector reaches."""
import pytest

import helpers

CASES = sorted(helpers.synth_manifest().keys())


@pytest.mark.parametrize("name", CASES)
def test_synth_is_deterministic_and_oracle_matches_reference(built, name):
    m = helpers.synth_manifest()[name]
    ivf = helpers.synth_stream(m["args"])
    assert helpers.md5(ivf) == m["ivf_md5"], "vp8synth output changed: regenerate the goldens"
    frames = helpers.oracle_decode_ivf(ivf)
    gold = helpers.synth_golden(name)
    assert len(frames) == len(gold) == m["shown_frames"]
    for k, (img, md5) in enumerate(zip(frames, gold)):
        assert helpers.md5(img) == md5, f"{name} frame {k}"

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


def test_randomised_streams_oracle_equals_reference(built, tmp_path):
    """The same kind of randomised generator settings as the GPU test, here oracle vs the compiled
    reference decoder when it is available (container only; the GPU box has no /root/reference but
    does carry oracle/_ref/decode)."""
    import os
    import random
    if not os.path.exists(helpers.REF_DECODE):
        pytest.skip("oracle/_ref/decode not built")
    rng = random.Random(7122)
    for k in range(12):
        w, h = rng.choice([(16, 16), (17, 33), (96, 80), (130, 98), (176, 144)])
        cfg = (f"--width {w} --height {h} --frames {rng.randint(3, 8)} --seed {500 + k} --version {rng.randint(0, 3)} "
               f"--filter-type {rng.randint(0, 1)} --sharpness {rng.randint(0, 7)} --lf {rng.choice([0, 8, 24, 63])} "
               f"--q {rng.choice([0, 40, 127])} --log2-parts {rng.randint(0, 3)} --key-interval {rng.choice([0, 3])} "
               f"--pct-intra {rng.choice([0, 8, 50])} --pct-split {rng.choice([0, 10, 60])} --pct-bpred {rng.choice([0, 25, 100])} "
               f"--golden-period {rng.choice([0, 2, 5])} --altref-period {rng.choice([0, 3])}")
        ivf = helpers.synth_stream(cfg)
        p = tmp_path / f"r{k}.ivf"
        p.write_bytes(ivf)
        ref = helpers.ref_decode_ivf(str(p))
        mine = b"".join(helpers.oracle_decode_ivf(ivf))
        assert mine == ref, cfg

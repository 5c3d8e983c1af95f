"""CPU tier: host parser -> oracle against every golden vector the reference ships
(example/vp8-test-vectors/*.ivf.md5, copied to tests/golden/vp8-test-vectors/).  This is what
pins the oracle (and the product's host parser) to the reference."""
import os

import pytest

import helpers


@pytest.mark.parametrize("ivf", helpers.vectors(), ids=lambda p: os.path.basename(p)[:-4])
def test_parser_plus_oracle_matches_golden_md5(built, ivf):
    frames = helpers.oracle_decode_ivf(ivf)
    gold = helpers.golden_md5(ivf)
    assert len(frames) == len(gold)
    for k, (img, (md5, w, h)) in enumerate(zip(frames, gold)):
        assert len(img) == helpers.i420_bytes(w, h), f"frame {k}: size"
        assert helpers.md5(img) == md5, f"frame {k}: md5"


def test_vector_inventory():
    # 43 streams, 700 shown frames
    vecs = helpers.vectors()
    assert len(vecs) == 43
    assert sum(len(helpers.golden_md5(v)) for v in vecs) == 700

"""Inputs the reference leaves undefined: the oracle reports them instead of following the reference into undefined
behaviour, the same way the C ABI does (ORB_E_ARG, orb_extract.cu: "aspect ratio < 0.5 is undefined in the reference")."""
import numpy as np
import pytest

from orb_slam3_b200.synth import synth_frame


def test_tall_images_are_reported_not_followed(oracle):
    """ORBextractor.cc:560-561: nIni = round(width / height) of the level's band; 0 for a band more than twice as tall as
    wide, then hX = width / 0 and vpIniNodes[...] of an empty vector."""
    ex = oracle.OracleExtractor(1000)
    with pytest.raises(ValueError):
        ex.extract(synth_frame(747, 305, 1))
    with pytest.raises(ValueError):
        ex.extract(synth_frame(747, 400, 2))      # level 0 is fine (nIni 1), a coarser level is not
    k, d, mono = ex.extract(synth_frame(700, 547, 3))   # the same object keeps working afterwards
    assert len(k) > 500 and mono == len(k)

"""GPU parity of bow_transform (C ABI; DBoW2 TemplatedVocabulary::transform as Frame::ComputeBoW calls it)
against the std::map oracle: BowVector ids and L1-normalised double weights bit for bit, FeatureVector
node lists and feature order exactly, from host descriptors and from an extractor's device results."""
import numpy as np
import pytest

from orb_slam3_b200 import scenes
from test_bow_oracle import KEYS, _descriptors

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k,L,levelsup,n", [(10, 4, 2, 2000), (8, 3, 4, 777), (10, 5, 4, 1200), (3, 6, 3, 2),
                                             (10, 4, 2, 2048), (10, 4, 2, 3000),
                                             (10, 5, 4, 8192), (10, 5, 4, 10000)])  # 10000: 5 x nFeatures at monocular initialisation
def test_host_descriptors_bitwise(oracle, k, L, levelsup, n):
    from orb_slam3_b200.bow import ORBVocabulary
    voc = scenes.synth_vocabulary(k, L, seed=7)
    desc = _descriptors(voc, n, seed=n)
    ref = oracle.bow_transform(voc, desc, levelsup)
    gv = ORBVocabulary(voc)
    got = gv.transform(desc, levelsup)
    assert got["used"] == ref["used"]
    for key in KEYS:
        assert np.array_equal(got[key], ref[key]), key
    again = gv.transform(desc, levelsup)                       # handle reuse, deterministic
    for key in KEYS:
        assert np.array_equal(again[key], ref[key]), key
    assert gv.kernel_launches() == 4 and gv.last_ms() > 0


def test_descriptors_taken_from_the_extractor_on_the_device(oracle):
    from orb_slam3_b200.bow import ORBVocabulary
    from orb_slam3_b200.extractor import ORBextractor
    from orb_slam3_b200.synth import synth_frame
    voc = scenes.synth_vocabulary(10, 4, seed=2)
    gv = ORBVocabulary(voc)
    ext = ORBextractor(1000, 1.2, 8, 20, 7)
    imgs = [synth_frame(480, 640, 3), synth_frame(480, 640, 4)]
    ext.extract_batch(imgs)
    for f, img in enumerate(imgs):
        _, d, _ = oracle.OracleExtractor(1000).extract(img)
        ref = oracle.bow_transform(voc, d, 2)
        got = gv.transform_extracted(ext, frame=f, levelsup=2)
        assert got["used"] == ref["used"] > 500
        for key in KEYS:
            assert np.array_equal(got[key], ref[key]), (f, key)
    e = gv.transform(np.zeros((0, 32), np.uint8), 2)
    assert e["used"] == 0 and list(e["fv_ptr"]) == [0]


def test_bad_arguments():
    from orb_slam3_b200._lib import OrbError
    from orb_slam3_b200.bow import ORBVocabulary
    from orb_slam3_b200.extractor import ORBextractor
    voc = scenes.synth_vocabulary(4, 2, seed=1)
    gv = ORBVocabulary(voc)
    with pytest.raises(OrbError):                                # extractor without results
        gv.transform_extracted(ORBextractor(500, 1.2, 8, 20, 7))
    voc.n_nodes = 1
    with pytest.raises(OrbError):
        ORBVocabulary(voc)

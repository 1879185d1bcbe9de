"""GPU parity of stereo_match (C ABI, Frame::ComputeStereoMatches Frame.cc:811-981) against the CPU
oracle: mvuRight and mvDepth bit-exact (float equality) on seeded rectified pairs, single and batched."""
import numpy as np
import pytest

from orb_slam3_b200.synth import synth_frame, stereo_right

pytestmark = pytest.mark.gpu

BF, B = 386.0, 0.5514


def _oracle_pair(oracle, left, right, nf):
    el, er = oracle.OracleExtractor(nf), oracle.OracleExtractor(nf)
    kl, dl, _ = el.extract(left)
    kr, dr, _ = er.extract(right)
    pl = [el.level_image(l) for l in range(8)]
    pr = [er.level_image(l) for l in range(8)]
    return kl, oracle.stereo_match(kl, dl, kr, dr, pl, pr, BF, B)


@pytest.fixture(scope="module")
def rig():
    from orb_slam3_b200.extractor import ORBextractor
    from orb_slam3_b200.stereo import StereoMatcher
    made = {}

    def get(nf):
        if nf not in made:
            made[nf] = (ORBextractor(nf, 1.2, 8, 20, 7), ORBextractor(nf, 1.2, 8, 20, 7), StereoMatcher())
        return made[nf]
    return get


@pytest.mark.parametrize("h,w,nf", [(480, 640, 1000), (720, 1280, 2000)])
@pytest.mark.parametrize("disp,noise", [((12,), 3), ((5, 30, 17), 3), ((0,), 2), ((40, 3), 0)])
def test_single_pair_bit_exact(oracle, rig, h, w, nf, disp, noise):
    left = synth_frame(h, w, 5 + len(disp))
    right = stereo_right(left, 100 + sum(disp), disparities=disp, noise=noise)
    kl, (n_ref, ur_ref, dp_ref, _) = _oracle_pair(oracle, left, right, nf)
    el, er, sm = rig(nf)
    _, gk, _ = el(left)
    er(right)
    assert np.array_equal(gk["x"], kl["x"]) and np.array_equal(gk["y"], kl["y"])
    n, ur, dp = sm.ComputeStereoMatches(el, er, len(gk), BF, B)
    assert n == n_ref, (n, n_ref)
    bad = np.nonzero(ur != ur_ref)[0]
    assert len(bad) == 0, (len(bad), bad[:5], ur[bad[:5]], ur_ref[bad[:5]])
    assert np.array_equal(dp, dp_ref)
    if noise:
        assert n > 100


def test_batch_of_pairs_and_state_reuse(oracle, rig):
    el, er, sm = rig(1000)
    lefts = [synth_frame(480, 640, 20 + i) for i in range(5)]
    rights = [stereo_right(l, 40 + i, disparities=(8 + 3 * i,)) for i, l in enumerate(lefts)]
    launches0 = sm.kernel_launches()
    res = el.extract_batch(lefts)
    er.extract_batch(rights)
    kept, ur, dp = sm.compute_batch(el, er, 5, BF, B)
    sm.compute_batch(el, er, 5, BF, B, on_device=True)      # enqueue-only mode leaves the same device state
    kept2, ur2, dp2 = sm.compute_batch(el, er, 5, BF, B)
    for i in range(5):
        kl, (n_ref, ur_ref, dp_ref, _) = _oracle_pair(oracle, lefts[i], rights[i], 1000)
        n = len(res[i][1])
        assert n == len(kl) and kept[i] == n_ref == kept2[i]
        assert np.array_equal(ur[i, :n], ur_ref) and np.array_equal(dp[i, :n], dp_ref)
        assert np.array_equal(ur2[i, :n], ur_ref) and np.array_equal(dp2[i, :n], dp_ref)
    assert sm.kernel_launches() - launches0 == 9 and sm.last_ms() > 0


def test_errors(rig):
    from orb_slam3_b200._lib import OrbError
    from orb_slam3_b200.extractor import ORBextractor
    from orb_slam3_b200.stereo import StereoMatcher
    el, er, sm = rig(1000)
    img = synth_frame(480, 640, 1)
    el(img)
    fresh = ORBextractor(1000, 1.2, 8, 20, 7)
    with pytest.raises(OrbError):                      # right handle has not extracted anything
        sm.ComputeStereoMatches(el, fresh, 10, BF, B)
    other = ORBextractor(500, 1.2, 8, 20, 7)
    other(img)
    with pytest.raises(OrbError):                      # different extractor parameters
        sm.ComputeStereoMatches(el, other, 10, BF, B)
    er(img)
    with pytest.raises(OrbError):                      # mb must be positive (maxD = mbf / mb)
        StereoMatcher().ComputeStereoMatches(el, er, 10, BF, 0.0)

"""The matcher / frustum / stereo oracle pinned to the reference's own object code.

oracle/_ref/libref_front.so is src/ORBmatcher.cc, src/Frame.cc, src/MapPoint.cc, src/ORBextractor.cc and
src/CameraModels/Pinhole.cpp of the reference compiled UNMODIFIED (oracle/Makefile) against the functional stand-ins of
oracle/cvcompat (cv::Mat, KeyPoint) and oracle/eigencompat (fixed-size Eigen, Sophus SE3).  The wrapper
(oracle/ref_front_wrap.cpp) fills real Frame / MapPoint objects from the flat views and calls

  ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th, bFarPoints, thFarPoints)   ORBmatcher.cc:40-285   (a10)
  ORBmatcher::SearchByProjection(Frame& Cur, const Frame& Last, th, bMono)                  ORBmatcher.cc:1669-1890 (a13)
  Frame::isInFrustum(MapPoint*, viewingCosLimit)                                           Frame.cc:542-612        (f-3)
  Frame::ComputeStereoMatches()                                                            Frame.cc:811-981        (f-1)

The restatements of oracle/orc_match.cpp, orc_frustum.cpp and orc_stereo.cpp have to give the SAME assignments and bit-equal
floats.  Built only where /root/reference exists (here); the prebuilt library travels to the GPU box."""
import numpy as np
import pytest

from orb_slam3_b200 import scenes
from orb_slam3_b200.synth import synth_frame, shifted_frame, stereo_right
from oracle import ref

pytestmark = pytest.mark.skipif(not ref.front_available(), reason="oracle/_ref/libref_front.so not built and no /root/reference")


@pytest.fixture(scope="module")
def feats(oracle):
    a = synth_frame(480, 640, 31)
    b = shifted_frame(a, 5, -3, 32)
    ex = oracle.OracleExtractor(1000)
    ka, da, _ = ex.extract(a)
    kb, db, _ = ex.extract(b)
    return ka, da, kb, db


@pytest.mark.parametrize("seed,th,ratio,stereo", [(5, 3.0, 0.8, False), (6, 1.0, 0.8, False), (7, 5.0, 0.9, True),
                                                  (8, 15.0, 0.6, True)])
def test_search_local_points_is_the_reference(oracle, feats, seed, th, ratio, stereo):
    ka, da, _, _ = feats
    F, mps = scenes.local_map_scene(ka, da, 640, 480, 300, seed=seed, stereo=stereo)
    n0, a0 = oracle.match_project_local(F, mps, th, ratio)
    n1, a1 = ref.front_project_local(F, mps, th, ratio)
    assert n0 == n1 and n0 > 50
    assert np.array_equal(a0, a1)


def test_search_local_points_far_points(oracle, feats):
    ka, da, _, _ = feats
    F, mps = scenes.local_map_scene(ka, da, 640, 480, 300, seed=11)
    th_far = 30.0   # the scene's predicted depths are uniform in [2, 80]
    n0, a0 = oracle.match_project_local(F, mps, 3.0, 0.8, far_points=True, th_far=th_far)
    n1, a1 = ref.front_project_local(F, mps, 3.0, 0.8, far_points=True, th_far=th_far)
    assert n0 == n1
    assert np.array_equal(a0, a1)
    n2, _ = ref.front_project_local(F, mps, 3.0, 0.8)
    assert 0 < n1 < n2


@pytest.mark.parametrize("seed,th,stereo,fwd,bwd,ori", [(2, 15.0, False, False, False, True), (3, 7.0, False, False, False, False),
                                                        (4, 15.0, True, False, False, True), (5, 15.0, True, True, False, True),
                                                        (6, 15.0, True, False, True, True), (9, 30.0, False, False, False, True)])
def test_search_last_frame_is_the_reference(oracle, feats, seed, th, stereo, fwd, bwd, ori):
    ka, da, kb, db = feats
    cur, last, Tcw = scenes.last_frame_scene(ka, da, kb, db, 640, 480, (5, -3), seed=seed, stereo=stereo)
    n0, a0 = oracle.match_project_last(cur, last, Tcw, th, forward=fwd, backward=bwd, check_ori=ori)
    n1, a1 = ref.front_project_last(cur, last, Tcw, th, forward=fwd, backward=bwd, check_ori=ori)
    assert n0 == n1 and n0 > 50
    # the reference leaves NULL both where nothing matched and where the rotation check cleared the match; the oracle
    # tells the two apart (-1 / -2)
    assert np.array_equal(np.where(a0 < 0, -1, a0), a1)


@pytest.mark.parametrize("seed,n", [(2, 3000), (3, 1), (4, 257)])
def test_is_in_frustum_is_the_reference(oracle, seed, n):
    v, _ = scenes.frustum_scene(n, seed=seed)
    k0, o0 = oracle.is_in_frustum(v, 0.5)
    k1, o1 = ref.front_is_in_frustum(v, 0.5)
    assert k0 == k1
    assert np.array_equal(o0["track_in_view"], o1["track_in_view"])
    for f in ("proj_x", "proj_y"):       # written for every point that reaches the projection (Frame.cc:566-567)
        assert np.array_equal(o0[f], o1[f]), f
    inside = o0["track_in_view"] != 0
    for f in ("proj_xr", "scale_level", "view_cos", "depth"):
        assert np.array_equal(o0[f][inside], o1[f][inside]), f


@pytest.mark.parametrize("seed,nf", [(41, 1000), (43, 400)])
def test_compute_stereo_matches_is_the_reference(oracle, seed, nf):
    left = synth_frame(480, 752, seed)
    right = stereo_right(left, seed + 1)
    el, er = oracle.OracleExtractor(nf), oracle.OracleExtractor(nf)
    kl, dl, _ = el.extract(left)
    kr, dr, _ = er.extract(right)
    pl = [el.level_image(l) for l in range(8)]
    pr = [er.level_image(l) for l in range(8)]
    n0, ur0, dp0, _ = oracle.stereo_match(kl, dl, kr, dr, pl, pr, 386.0, 0.5514)
    n1, ur1, dp1 = ref.front_stereo_match(kl, dl, kr, dr, pl, pr, 386.0, 0.5514)
    assert n0 == n1 and n0 > 20
    assert np.array_equal(ur0, ur1)
    assert np.array_equal(dp0, dp1)


def _qt(axis_angle, t):
    a = np.asarray(axis_angle, np.float64)
    th = np.linalg.norm(a)
    q = np.array([0, 0, 0, 1.0]) if th == 0 else np.concatenate([a / th * np.sin(th / 2), [np.cos(th / 2)]])
    return np.concatenate([q, t]).astype(np.float32)


@pytest.mark.parametrize("seed,stereo", [(3, True), (4, False), (5, True)])
@pytest.mark.parametrize("pose", ["along_shift", "rotated", "across_shift"])
def test_search_for_triangulation_is_the_reference(oracle, feats, seed, stereo, pose):
    """SearchForTriangulation (ORBmatcher.cc:907-1146, a11) on two real KeyFrame objects.  The oracle (like the C ABI) takes
    F12 and the epipole from its caller; here they are the ones the reference derives from the two poses."""
    ka, da, kb, db = feats
    k1, k2, fv1, fv2, _, _ = scenes.triangulation_scene(ka, da, kb, db, 640, 480, seed=seed, stereo=stereo)
    T1 = _qt([0, 0, 0], [0.3, -0.1, 0.2])
    t = {"along_shift": [-0.1, 0.06, -0.003], "rotated": [-0.1, 0.06, -0.003], "across_shift": [0.05, 0.09, 0.004]}[pose]
    T2 = _qt([0.002, -0.003, 0.004] if pose == "rotated" else [0, 0, 0], np.add(T1[4:], t))
    counts = []
    for only_stereo, coarse, ori in [(0, 0, 1), (1, 0, 1), (0, 1, 1), (0, 0, 0)]:
        n1, p1, F12, ep = ref.front_triangulate(k1, k2, fv1, fv2, T1, T2, only_stereo, coarse, ori)
        n0, p0 = oracle.match_triangulate(k1, k2, fv1, fv2, F12, ep, only_stereo, coarse, ori)
        assert n0 == n1
        assert np.array_equal(p0, p1)
        counts.append(n1)
    assert counts[2] > 50                  # the coarse pass ignores the geometry
    assert counts[0] <= counts[2] + 3      # (rotation-histogram pruning can differ by a few between the passes)
    if pose == "across_shift":
        assert counts[0] < counts[2] // 2  # an epipolar geometry the true shift contradicts rejects most candidates


def test_matchers_degenerate_inputs_follow_the_reference(oracle, feats):
    """Nothing in view; identical descriptors everywhere (every decision is a tie, resolved by the order in which
    GetFeaturesInArea returns the candidates); map points without observations (never blocking, slots overwritten);
    an empty map-point list."""
    ka, da, kb, db = feats
    F, mps = scenes.local_map_scene(ka, da, 640, 480, 10, seed=1)
    mps._keep["track_in_view"][:] = 0
    n0, a0 = oracle.match_project_local(F, mps, 3.0, 0.8)
    n1, a1 = ref.front_project_local(F, mps, 3.0, 0.8)
    assert n0 == n1 == 0 and np.array_equal(a0, a1)

    d0 = np.zeros_like(da)
    F, mps = scenes.local_map_scene(ka, d0, 640, 480, 0, seed=2)
    mps._keep["desc"][:] = 0
    for th in (1.0, 3.0, 10.0):
        n0, a0 = oracle.match_project_local(F, mps, th, 0.8)
        n1, a1 = ref.front_project_local(F, mps, th, 0.8)
        assert n0 == n1 and np.array_equal(a0, a1), th

    cur, last, Tcw = scenes.last_frame_scene(ka, da, kb, db, 640, 480, (5, -3), seed=4, obs0_frac=1.0)
    n0, a0 = oracle.match_project_last(cur, last, Tcw, 30.0, check_ori=True)
    n1, a1 = ref.front_project_last(cur, last, Tcw, 30.0, check_ori=True)
    assert n0 == n1 and np.array_equal(np.where(a0 < 0, -1, a0), a1)

    cur, last, Tcw = scenes.last_frame_scene(ka, np.zeros_like(da), kb, np.zeros_like(db), 640, 480, (5, -3), seed=5)
    last._keep["desc"][:] = 0
    n0, a0 = oracle.match_project_last(cur, last, Tcw, 15.0, check_ori=False)
    n1, a1 = ref.front_project_last(cur, last, Tcw, 15.0, check_ori=False)
    assert n0 == n1 and np.array_equal(np.where(a0 < 0, -1, a0), a1)

    F, mps = scenes.local_map_scene(ka[:0], da[:0], 640, 480, 0, seed=3)
    assert F.n == 0 and mps.n == 0
    n0, a0 = oracle.match_project_local(F, mps, 3.0, 0.8)
    n1, a1 = ref.front_project_local(F, mps, 3.0, 0.8)
    assert n0 == n1 == 0


def test_compute_stereo_matches_degenerate_inputs_follow_the_reference(oracle):
    """No right keypoints; identical images (zero disparity, every SAD window a perfect match: the sub-pixel parabola
    degenerates); several disparity bands with noise-free windows."""
    left = synth_frame(480, 640, 12)
    el = oracle.OracleExtractor(800)
    kl, dl, _ = el.extract(left)
    pl = [el.level_image(l) for l in range(8)]
    for right in (left.copy(), stereo_right(left, 13, disparities=(40, 3), noise=0)):
        er = oracle.OracleExtractor(800)
        kr, dr, _ = er.extract(right)
        pr = [er.level_image(l) for l in range(8)]
        n0, ur0, dp0, _ = oracle.stereo_match(kl, dl, kr, dr, pl, pr, 386.0, 0.5514)
        n1, ur1, dp1 = ref.front_stereo_match(kl, dl, kr, dr, pl, pr, 386.0, 0.5514)
        assert n0 == n1 and np.array_equal(ur0, ur1) and np.array_equal(dp0, dp1)
    n0, ur0, dp0, _ = oracle.stereo_match(kl, dl, kr[:0], dr[:0], pl, pr, 386.0, 0.5514)
    n1, ur1, dp1 = ref.front_stereo_match(kl, dl, kr[:0], dr[:0], pl, pr, 386.0, 0.5514)
    assert n0 == n1 == 0 and np.array_equal(ur0, ur1) and np.array_equal(dp0, dp1)

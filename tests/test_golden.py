"""Committed golden vectors (tests/golden/, made by scripts/make_golden.py in the build
container): cv2's own outputs pin the oracle's OpenCV primitives where cv2 is absent; the oracle's
outputs on seeded inputs pin both the oracle (regression) and the GPU path."""
import os

import numpy as np
import pytest

from orb_slam3_b200 import scenes
from orb_slam3_b200.synth import synth_frame, shifted_frame

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(G, name), allow_pickle=False)


def test_oracle_primitives_match_cv2_golden(oracle):
    z = _load("primitives_cv2.npz")
    assert np.array_equal(oracle.resize_linear_u8(z["img"], 133, 100), z["resize_img"])
    assert np.array_equal(oracle.resize_linear_u8(z["rnd"], 109, 81), z["resize_rnd"])
    assert np.array_equal(oracle.gaussian_blur7(z["img"]), z["blur_img"])
    assert np.array_equal(oracle.gaussian_blur7(z["rnd"]), z["blur_rnd"])
    assert np.array_equal(oracle.fast(z["img"], 20), z["fast20"])
    assert np.array_equal(oracle.fast(z["img"], 7), z["fast7"])
    got = np.array([oracle.fast_atan2(y, x) for y, x in zip(z["atan_y"], z["atan_x"])], np.float32)
    assert np.array_equal(got, z["atan"])
    assert np.array_equal(synth_frame(120, 160, 3), z["img"])  # the generator itself is pinned too


def _scene():
    z = _load("extract_640x480.npz")
    f0 = synth_frame(480, 640, 1)
    f1 = shifted_frame(f0, 5, -3, 2)
    return z, f0, f1


def test_oracle_reproduces_golden(oracle):
    z, f0, f1 = _scene()
    k, d, mono = oracle.OracleExtractor(1000).extract(f0)
    assert mono == int(z["mono"]) and np.array_equal(d, z["desc"])
    for f in k.dtype.names:
        assert np.array_equal(k[f], z["kps"][f]), f
    k1, d1, _ = oracle.OracleExtractor(1000).extract(f1)
    m = _load("match_scene.npz")
    cur, last, Tcw = scenes.last_frame_scene(k, d, k1, d1, 640, 480, (5, -3), seed=3, stereo=True)
    n, a = oracle.match_project_last(cur, last, Tcw, 15.0)
    assert n == int(m["n_last"]) and np.array_equal(a, m["a_last"])
    F, mps = scenes.local_map_scene(k1, d1, 640, 480, 400, seed=4)
    n, a = oracle.match_project_local(F, mps, 3.0, 0.8)
    assert n == int(m["n_loc"]) and np.array_equal(a, m["a_loc"])
    kf1, kf2, fv1, fv2, F12, ep = scenes.triangulation_scene(k, d, k1, d1, 640, 480, seed=5, n_nodes=60)
    n, p = oracle.match_triangulate(kf1, kf2, fv1, fv2, F12, ep)
    assert n == int(m["n_tri"]) and np.array_equal(p, m["pairs"])
    g, _ = scenes.lba_graph(8, 300, seed=1)
    r = oracle.lba_solve(scenes.lba_view(g))
    zl = _load("lba_small.npz")
    assert r["iterations"] == int(zl["iterations"]) and r["stats"]["trials"] == int(zl["trials"])
    assert np.allclose(r["kf_pose"], zl["kf_pose"], rtol=0, atol=1e-9)
    assert np.allclose(r["mp_pos"], zl["mp_pos"], rtol=0, atol=1e-7)
    # fisheye stereo rig (row a17): KannalaBrandt8 mono edges + second-camera edges
    g, _ = scenes.lba_rig_graph(8, 300, seed=1)
    assert int((g["e_stereo"] == 2).sum()) == int(_load("lba_rig_small.npz")["n_body"])
    r = oracle.lba_solve(scenes.lba_view(g))
    zr = _load("lba_rig_small.npz")
    assert r["iterations"] == int(zr["iterations"]) and r["stats"]["trials"] == int(zr["trials"])
    assert np.allclose(r["kf_pose"], zr["kf_pose"], rtol=0, atol=1e-9)
    assert np.allclose(r["mp_pos"], zr["mp_pos"], rtol=0, atol=1e-7)


def _stereo_inputs():
    from orb_slam3_b200.synth import stereo_right
    sl = synth_frame(480, 640, 5)
    return sl, stereo_right(sl, 6, disparities=(5, 30, 17))


def test_oracle_reproduces_stereo_and_pose_golden(oracle):
    sl, sr = _stereo_inputs()
    el, er = oracle.OracleExtractor(1000), oracle.OracleExtractor(1000)
    kl, dl, _ = el.extract(sl)
    kr, dr, _ = er.extract(sr)
    n, ur, dp, sad = oracle.stereo_match(kl, dl, kr, dr, [el.level_image(l) for l in range(8)],
                                         [er.level_image(l) for l in range(8)], 386.0, 0.5514)
    z = _load("stereo_640x480.npz")
    assert n == int(z["n"]) and np.array_equal(ur, z["u_right"]) and np.array_equal(dp, z["depth"])
    assert np.array_equal(sad, z["sad"])
    pv, _ = scenes.pose_scene(400, seed=7)
    r = oracle.pose_optimize(pv)
    zp = _load("pose_small.npz")
    assert r["inliers"] == int(zp["inliers"]) and np.array_equal(r["outlier"], zp["outlier"])
    assert np.array_equal(r["stats"], zp["stats"]) and np.allclose(r["pose"], zp["pose"], rtol=0, atol=1e-11)


def test_oracle_reproduces_frustum_bow_and_lia_golden(oracle):
    fv, _ = scenes.frustum_scene(3000, seed=2)
    n_in, fo = oracle.is_in_frustum(fv, 0.5)
    z = _load("frustum_small.npz")
    assert n_in == int(z["n_in"])
    for k in fo:
        assert np.array_equal(fo[k], z[k]), k
    zs, f0, _ = _scene()
    voc = scenes.synth_vocabulary(10, 4, seed=2)
    bw = oracle.bow_transform(voc, zs["desc"], 2)
    zb = _load("bow_small.npz")
    for k in ("bow_ids", "bow_vals", "fv_node_ids", "fv_ptr", "fv_idx"):
        assert np.array_equal(bw[k], zb[k]), k
    ld, _ = scenes.lia_scene(5, 150, seed=6)
    lr = oracle.lia_solve(oracle.make_lia_view(ld))
    zl = _load("lia_small.npz")
    assert lr["stats"]["iterations"] == int(zl["iterations"]) and lr["stats"]["trials"] == int(zl["trials"])
    for k in ("tcw", "Rcw", "vel", "bg", "ba", "mp_pos"):
        assert np.allclose(lr[k], zl[k], rtol=0, atol=1e-9), k
    assert np.allclose(lr["chi2"], zl["chi2"], rtol=1e-9, atol=1e-9)


@pytest.mark.gpu
def test_gpu_reproduces_frustum_and_bow_golden():
    from orb_slam3_b200.bow import ORBVocabulary
    from orb_slam3_b200.frustum import FrustumCuller
    fv, _ = scenes.frustum_scene(3000, seed=2)
    n_in, fo = FrustumCuller().isInFrustum(fv, 0.5)
    z = _load("frustum_small.npz")
    assert n_in == int(z["n_in"])
    for k in fo:
        assert np.array_equal(fo[k], z[k]), k
    zs = _load("extract_640x480.npz")
    bw = ORBVocabulary(scenes.synth_vocabulary(10, 4, seed=2)).transform(zs["desc"], 2)
    zb = _load("bow_small.npz")
    for k in ("bow_ids", "bow_vals", "fv_node_ids", "fv_ptr", "fv_idx"):
        assert np.array_equal(bw[k], zb[k]), k


@pytest.mark.gpu
def test_gpu_reproduces_stereo_and_pose_golden():
    from orb_slam3_b200.extractor import ORBextractor
    from orb_slam3_b200.optimizer import PoseOptimization
    from orb_slam3_b200.stereo import StereoMatcher
    sl, sr = _stereo_inputs()
    el, er = ORBextractor(1000, 1.2, 8, 20, 7), ORBextractor(1000, 1.2, 8, 20, 7)
    _, kl, _ = el(sl)
    er(sr)
    n, ur, dp = StereoMatcher().ComputeStereoMatches(el, er, len(kl), 386.0, 0.5514)
    z = _load("stereo_640x480.npz")
    assert n == int(z["n"]) and np.array_equal(ur, z["u_right"]) and np.array_equal(dp, z["depth"])
    pv, _ = scenes.pose_scene(400, seed=7)
    inl, pose, out = PoseOptimization()(pv)
    zp = _load("pose_small.npz")
    assert inl == int(zp["inliers"]) and np.array_equal(out, zp["outlier"])
    assert np.allclose(pose, zp["pose"], rtol=0, atol=1e-8)


@pytest.mark.gpu
def test_gpu_reproduces_golden():
    from orb_slam3_b200.extractor import ORBextractor
    from orb_slam3_b200.matcher import ORBmatcher
    from orb_slam3_b200.optimizer import LocalBundleAdjustment
    z, f0, f1 = _scene()
    ext = ORBextractor(1000, 1.2, 8, 20, 7)
    mono, k, d = ext(f0)
    assert mono == int(z["mono"]) and np.array_equal(d, z["desc"])
    for f in k.dtype.names:
        assert np.array_equal(k[f], z["kps"][f]), f
    _, k1, d1 = ext(f1)
    m = _load("match_scene.npz")
    cur, last, Tcw = scenes.last_frame_scene(k, d, k1, d1, 640, 480, (5, -3), seed=3, stereo=True)
    n, a = ORBmatcher(0.9, True).SearchByProjectionLast(cur, last, Tcw, 15.0)
    assert n == int(m["n_last"]) and np.array_equal(a, m["a_last"])
    F, mps = scenes.local_map_scene(k1, d1, 640, 480, 400, seed=4)
    n, a = ORBmatcher(0.8).SearchByProjection(F, mps, 3.0)
    assert n == int(m["n_loc"]) and np.array_equal(a, m["a_loc"])
    kf1, kf2, fv1, fv2, F12, ep = scenes.triangulation_scene(k, d, k1, d1, 640, 480, seed=5, n_nodes=60)
    n, p = ORBmatcher(0.6, True).SearchForTriangulation(kf1, kf2, fv1, fv2, F12, ep)
    assert n == int(m["n_tri"]) and np.array_equal(p, m["pairs"])
    g, _ = scenes.lba_graph(8, 300, seed=1)
    r = LocalBundleAdjustment()(scenes.lba_view(g))
    zl = _load("lba_small.npz")
    assert r["iterations"] == int(zl["iterations"]) and r["stats"]["trials"] == int(zl["trials"])
    step = np.abs(zl["mp_pos"] - g["mp_pos"]).max()
    assert np.abs(r["mp_pos"] - zl["mp_pos"]).max() < 1e-4 * step
    assert np.abs(r["kf_pose"] - zl["kf_pose"]).max() < 1e-6
    g, _ = scenes.lba_rig_graph(8, 300, seed=1)
    r = LocalBundleAdjustment()(scenes.lba_view(g))
    zr = _load("lba_rig_small.npz")
    assert r["iterations"] == int(zr["iterations"]) and r["stats"]["trials"] == int(zr["trials"])
    step = np.abs(zr["mp_pos"] - g["mp_pos"]).max()
    assert np.abs(r["mp_pos"] - zr["mp_pos"]).max() < 1e-4 * step
    # poses: 1e-4 of the largest pose update (the fisheye window moves its keyframes by centimetres; its float atan2f is
    # evaluated differently on the device, see test_lba_gpu.py::test_lba_second_camera_and_fisheye_edges)
    pose_step = np.abs(zr["kf_pose"] - g["kf_pose"]).max()
    assert np.abs(r["kf_pose"] - zr["kf_pose"]).max() < 1e-4 * pose_step, pose_step

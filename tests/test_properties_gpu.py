"""Size-independent properties at BASELINE.json's full sizes (no oracle involved):
extractor invariants on 1280x720 / 2000 features in a batch, matcher invariants, LBA
monotonicity on the config-4 graph."""
import numpy as np
import pytest

from orb_slam3_b200 import scenes
from orb_slam3_b200.synth import synth_frame, shifted_frame

pytestmark = pytest.mark.gpu


def test_extractor_invariants_full_size_batch():
    from orb_slam3_b200.extractor import ORBextractor
    ext = ORBextractor(2000, 1.2, 8, 20, 7)
    base = synth_frame(720, 1280, 77)
    frames = [base, shifted_frame(base, 4, 2, 78), base, synth_frame(720, 1280, 79, low_texture=True)]
    res = ext.extract_batch(frames)
    quota = ext.features_per_level()
    sf = ext.GetScaleFactors()
    for mono, k, d in res:
        assert mono == len(k) and len(k) == len(d)
        assert 0 < len(k) <= 2000 + 3 * 8
        for lvl in range(8):
            sel = k[k["octave"] == lvl]
            assert len(sel) <= quota[lvl] + 3                       # DistributeOctTree overshoot bound
            w = int(np.rint(np.float32(1280) / sf[lvl]))
            h = int(np.rint(np.float32(720) / sf[lvl]))
            x, y = sel["x"] / sf[lvl], sel["y"] / sf[lvl]
            assert (x >= 19 - 1e-3).all() and (x <= w - 19 + 1e-3).all()  # EDGE_THRESHOLD
            assert (y >= 19 - 1e-3).all() and (y <= h - 19 + 1e-3).all()
            assert (sel["size"] == np.float32(int(31 * sf[lvl]))).all()
        assert ((k["angle"] >= 0) & (k["angle"] < 360)).all()
        assert (k["response"] >= 7).all() and (k["class_id"] == -1).all()
        # no duplicate keypoints
        assert len(np.unique(np.stack([k["octave"], k["x"], k["y"]], 1), axis=0)) == len(k)
    # identical frames in one batch -> identical results (no cross-frame interference)
    assert np.array_equal(res[0][1], res[2][1]) and np.array_equal(res[0][2], res[2][2])
    # and run-to-run determinism
    again = ext.extract_batch(frames)
    for a, b in zip(res, again):
        assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    # the host pyramid mirror: level 0 is the input
    assert np.array_equal(ext.image_pyramid(0, frame=1), frames[1])


def test_matcher_invariants_full_size():
    from orb_slam3_b200.extractor import ORBextractor
    from orb_slam3_b200.matcher import ORBmatcher
    ext = ORBextractor(2000, 1.2, 8, 20, 7)
    a = synth_frame(720, 1280, 5)
    b = shifted_frame(a, 6, -4, 6)
    (_, ka, da), (_, kb, db) = ext.extract_batch([a, b])
    cur, last, Tcw = scenes.last_frame_scene(ka, da, kb, db, 1280, 720, (6, -4), seed=2, obs0_frac=0.0)
    n, assign = ORBmatcher(0.9, False).SearchByProjectionLast(cur, last, Tcw, 15.0)
    hit = assign[assign >= 0]
    assert n == len(hit) and n > 500
    assert len(np.unique(hit)) == len(hit)                      # a map point is assigned at most once
    taken = cur._keep[5].astype(bool)
    assert not (assign[taken] >= 0).any()                       # slots held by observed points are never overwritten
    K = last._keep
    d_mp = K["desc"][hit]
    d_kp = db[np.nonzero(assign >= 0)[0]]
    dist = np.unpackbits(d_mp ^ d_kp, axis=1).sum(1)
    assert (dist <= 100).all()                                  # TH_HIGH
    # a larger window can only find at least as many
    n2, _ = ORBmatcher(0.9, False).SearchByProjectionLast(cur, last, Tcw, 30.0)
    assert n2 >= n * 0.9


def test_lba_monotone_and_robust_full_size():
    from orb_slam3_b200.optimizer import LocalBundleAdjustment
    g, truth = scenes.lba_graph(50, 20000, seed=3)
    r = LocalBundleAdjustment()(scenes.lba_view(g))
    st = r["stats"]
    assert st["iterations"] >= 1 and st["trials"] >= st["iterations"]
    assert st["chi2_final"] < 0.5 * st["chi2_initial"]
    assert np.isfinite(r["kf_pose"]).all() and np.isfinite(r["mp_pos"]).all()
    q = r["kf_pose"][:, :4]
    assert np.allclose(np.linalg.norm(q, axis=1), 1.0, atol=1e-12) and (q[:, 3] >= 0).all()
    out = LocalBundleAdjustment.outliers(g, r)
    assert 0.005 < out.mean() < 0.15                            # the 2 % planted gross outliers (+ tails)
    # poses move towards the truth
    e0 = np.abs(g["kf_pose"][:, 4:] - truth["kf_pose"][:, 4:]).mean()
    e1 = np.abs(r["kf_pose"][:, 4:] - truth["kf_pose"][:, 4:]).mean()
    assert e1 < e0

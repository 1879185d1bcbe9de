"""CPU pinning of the Optimizer::PoseOptimization oracle (oracle/orc_pose.cpp, Optimizer.cc:814-1115):
independent numpy checks of what the restatement must satisfy -- stationarity of the last round's
(non-robust, inliers-only) cost at the returned pose, recovery of the true pose and of the planted
outliers, the chi2 classification rule, and the early exits."""
import numpy as np
import pytest

from orb_slam3_b200 import scenes


def _arrays(v):
    k = v._keep
    return k["xw"].astype(np.float64), k["obs"].astype(np.float64), k["inv_sigma2"].astype(np.float64)


def _rot(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _residuals(v, pose):
    """obs - projection, written independently of the oracle (double invz everywhere)."""
    xw, obs, _ = _arrays(v)
    Xc = xw @ _rot(pose[:4]).T + pose[4:]
    pu = v.fx * Xc[:, 0] / Xc[:, 2] + v.cx
    pv = v.fy * Xc[:, 1] / Xc[:, 2] + v.cy
    st = obs[:, 2] >= 0
    r = np.stack([obs[:, 0] - pu, obs[:, 1] - pv, np.where(st, obs[:, 2] - (pu - v.bf / Xc[:, 2]), 0.0)], 1)
    return r, st


def _chi2(v, pose):
    r, _ = _residuals(v, pose)
    return (r * r).sum(1) * _arrays(v)[2]


def _perturb(pose, d):
    """exp(d) * pose for a small twist d = (omega, upsilon), first order in upsilon."""
    w = d[:3]
    th = np.linalg.norm(w)
    dq = np.concatenate([np.sin(th / 2) * w / max(th, 1e-300), [np.cos(th / 2)]])
    x1, y1, z1, w1 = dq
    x2, y2, z2, w2 = pose[:4]
    q = np.array([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
                  w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2, w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2])
    return np.concatenate([q, _rot(dq) @ pose[4:] + d[3:]])


@pytest.mark.parametrize("seed,stereo_frac", [(0, 0.8), (1, 0.0), (2, 1.0)])
def test_result_is_a_stationary_point_and_classifies_by_chi2(oracle, seed, stereo_frac):
    v, truth = scenes.pose_scene(800, seed=seed, stereo_frac=stereo_frac)
    r = oracle.pose_optimize(v)
    assert r["stats"][0] == 4 and r["inliers"] == 800 - r["outlier"].sum()
    # classification rule of the last round (:1032-1044, :1080-1092), float compare
    chi2 = _chi2(v, r["pose"])
    th = np.where(truth["stereo"], np.float32(7.815), np.float32(5.991))
    # the oracle keeps the reference's float invz in the stereo error (1e-4 px effects at u ~ 1e3 px)
    clear = np.abs(chi2 - th) > 1e-2
    assert np.array_equal((chi2.astype(np.float32) > th)[clear], r["outlier"][clear])
    assert np.allclose(r["chi2"], chi2, rtol=1e-4, atol=1e-3)
    # round 3 minimises the plain chi2 of round 2's inliers: gradient ~ 0 at the result.  The inlier set of
    # round 3 is not returned, but away from the threshold it equals the final classification.
    inl = ~r["outlier"]

    def cost(p):
        return _chi2(v, p)[inl].sum()
    c0 = cost(r["pose"])
    g = np.array([(cost(_perturb(r["pose"], h * e)) - cost(_perturb(r["pose"], -h * e))) / (2 * h)
                  for e in np.eye(6) for h in [1e-6]])
    H = np.array([(cost(_perturb(r["pose"], 1e-4 * e)) - 2 * c0 + cost(_perturb(r["pose"], -1e-4 * e))) / 1e-8
                  for e in np.eye(6)])
    assert (np.abs(g) / np.sqrt(H * max(c0, 1.0)) < 2e-2).all(), g   # LM stops by its 3-strike rule, not at 0
    # planted outliers found, true pose recovered within the noise
    assert (r["outlier"][truth["outlier"]]).mean() > 0.95
    assert (r["outlier"][~truth["outlier"]]).mean() < 0.1
    dq = abs(np.dot(r["pose"][:4], truth["pose"][:4]))
    assert 2 * np.arccos(min(dq, 1.0)) < np.deg2rad(0.1) and np.abs(r["pose"][4:] - truth["pose"][4:]).max() < 0.02


def test_jacobians_by_finite_differences_through_one_lm_step(oracle):
    """No outliers, tiny noise: LM must reach the least-squares optimum that a numpy Gauss-Newton on
    finite-difference Jacobians of the independent residual function reaches."""
    v, _ = scenes.pose_scene(300, seed=5, outlier_frac=0.0)
    r = oracle.pose_optimize(v)
    pose = r["pose"].copy()
    w = np.sqrt(_arrays(v)[2])[:, None]
    for _ in range(5):  # Gauss-Newton from the oracle's answer: the step must be ~0
        f0 = (_residuals(v, pose)[0] * w).ravel()
        J = np.stack([((_residuals(v, _perturb(pose, 1e-6 * e))[0] * w).ravel() - f0) / 1e-6 for e in np.eye(6)], 1)
        step = np.linalg.lstsq(J, -f0, rcond=None)[0]
        pose = _perturb(pose, step)
    assert np.abs(pose[4:] - r["pose"][4:]).max() < 2e-4 and np.abs(pose[:4] - r["pose"][:4]).max() < 2e-5


def test_early_exits_and_rejected_trials(oracle):
    v, _ = scenes.pose_scene(2, seed=1)
    r = oracle.pose_optimize(v)                     # < 3 correspondences (:1000-1001)
    assert r["inliers"] == 0 and r["stats"][0] == 0
    q = np.array(v.pose[:4]); q /= np.linalg.norm(q)
    assert np.allclose(r["pose"][:4], q) and np.allclose(r["pose"][4:], v.pose[4:7])
    v, _ = scenes.pose_scene(8, seed=2, outlier_frac=0.0)
    r = oracle.pose_optimize(v)                     # < 10 edges: one round only (:1098-1099)
    assert r["stats"][0] == 1 and r["inliers"] == 8 - r["outlier"].sum()
    v, truth = scenes.pose_scene(600, seed=3, wild=True)
    r = oracle.pose_optimize(v)                     # a far start: more trials than iterations
    assert r["stats"][2] > r["stats"][1] and r["stats"][0] == 4

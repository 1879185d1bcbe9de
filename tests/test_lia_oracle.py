"""CPU pinning of the Optimizer::LocalInertialBA oracle (oracle/orc_lia.cpp, Optimizer.cc:2383-2958 with the
vertex / edge types of G2oTypes.cc; SURVEY.md 8(f-4b) -- oracle only, the CUDA path is the next round's work):
the assembled right-hand side must be the numerical gradient of the robust cost under the vertices' own
oplus (this checks every Jacobian block of EdgeMono / EdgeStereo / EdgeInertial / EdgeGyroRW / EdgeAccRW and
the ImuCamPose update at once), and the solver must pull a perturbed visual-inertial window back to the
trajectory that generated its IMU samples."""
import numpy as np
import pytest

from orb_slam3_b200 import scenes


def test_rhs_is_the_numerical_gradient_of_the_robust_cost(oracle):
    d, _ = scenes.lia_scene(4, 60, seed=2)
    v = oracle.make_lia_view(d)
    chi, b, _, n_pose = oracle.lia_linearize(v)
    assert n_pose == 4 * 15 and chi > 0
    n = len(b)
    rng = np.random.default_rng(0)
    idx = list(range(n_pose)) + list(rng.choice(np.arange(n_pose, n), 30, replace=False))
    worst = 0.0
    for i in idx:
        # the bias -> preintegration path runs in float like the reference (ImuTypes.cc:298-332): its rounding
        # needs a larger central-difference step (the cost is quadratic in the accelerometer bias anyway)
        h = 2e-3 if (i < n_pose and i % 15 >= 9) else 1e-5
        dl = np.zeros(n)
        dl[i] = h
        cp = oracle.lia_linearize(v, dl)[2]
        dl[i] = -h
        cm = oracle.lia_linearize(v, dl)[2]
        fd = -(cp - cm) / (4 * h)                      # b = -1/2 d(robust chi2)/d(delta)
        worst = max(worst, abs(fd - b[i]) / max(abs(b[i]), 1e-4 * np.abs(b).max()))
    assert worst < 2e-2, worst


@pytest.mark.parametrize("seed,n_opt", [(1, 6), (3, 10)])
def test_recovers_the_generating_trajectory(oracle, seed, n_opt):
    d, truth = scenes.lia_scene(n_opt, 400, seed=seed)
    v = oracle.make_lia_view(d)
    r = oracle.lia_solve(v)
    st = r["stats"]
    assert 1 <= st["iterations"] <= 10 and st["trials"] >= st["iterations"] and st["dim"] == 15 * n_opt
    assert st["err_end"] < 0.2 * st["err"] and st["lambda_final"] > 0
    K = v.n_kf
    tcw_t = np.stack([truth["cam_of"](truth["Rwb"][k], truth["twb"][k])[1] for k in range(K)])
    Rcw_t = np.stack([truth["cam_of"](truth["Rwb"][k], truth["twb"][k])[0] for k in range(K)])
    free = np.arange(n_opt)
    e0, e1 = np.abs(d["kf_tcw"] - tcw_t)[free].max(), np.abs(r["tcw"] - tcw_t)[free].max()
    assert e1 < 0.5 * e0 and e1 < 0.015, (e0, e1)
    assert np.abs(r["Rcw"] - Rcw_t)[free].max() < np.abs(d["kf_Rcw"].reshape(-1, 3, 3) - Rcw_t)[free].max()
    v0, v1 = np.abs(d["kf_vel"] - truth["vel"])[free].max(), np.abs(r["vel"] - truth["vel"])[free].max()
    assert v1 < 0.5 * v0, (v0, v1)
    assert np.abs(r["bg"][free] - truth["bg"]).max() < 2e-3 and np.abs(r["ba"][free] - truth["ba"]).max() < 5e-2
    # fixed keyframes are untouched
    assert np.array_equal(r["tcw"][n_opt:], d["kf_tcw"][n_opt:]) and np.array_equal(r["vel"][n_opt:], d["kf_vel"][n_opt:])
    # points: the 1..3.6 px observation noise dominates far away; near the cameras they stay centimetre-accurate
    near = (truth["mp_pos"][:, 0] - truth["twb"][n_opt][0]) < 10.0
    assert np.median(np.abs(r["mp_pos"] - truth["mp_pos"])[near]) < 0.03
    # planted gross observation errors end up above the chi2 gates (Optimizer.cc:2760-2790)
    st_e = np.asarray(d["e_stereo"], bool)
    flagged = r["chi2"] > np.where(st_e, 7.815, 5.991)
    assert 0.01 < flagged.mean() < 0.12 and r["depth_pos"].all()


def test_last_inertial_edge_is_downweighted(oracle):
    """i == N-1: information * 1e-2 and a Huber kernel (Optimizer.cc:2585-2596) -- turning the flag off
    changes the cost at the start by the expected factor on that edge only."""
    d, _ = scenes.lia_scene(3, 40, seed=4)
    c_on = oracle.lia_linearize(oracle.make_lia_view(d))[0]
    d2 = dict(d)
    d2["i_last"] = np.zeros(3, np.uint8)
    c_off = oracle.lia_linearize(oracle.make_lia_view(d2))[0]
    assert c_off > c_on

"""The product's LocalInertialBA source (csrc/lia_core.h, the body of lia_kernel) executed single-threaded on
the host through lia_debug_host must agree with the independent oracle (oracle/orc_lia.cpp): same LM
iteration / trial counts, identical chi2 classification inputs, states equal to rounding.  This is what
stands in for the GPU parity run of row 8(f-4b) until the kernel has been on hardware."""
import numpy as np
import pytest

from orb_slam3_b200 import scenes


def _compare(d, ref, got, tol=1e-7, err_tol=1e-9):
    assert got["stats"]["iterations"] == ref["stats"]["iterations"] and got["stats"]["trials"] == ref["stats"]["trials"]
    assert got["stats"]["dim"] == ref["stats"]["dim"]
    assert abs(got["stats"]["err"] - ref["stats"]["err"]) <= err_tol * ref["stats"]["err"]
    assert abs(got["stats"]["err_end"] - ref["stats"]["err_end"]) <= 1e-6 * ref["stats"]["err_end"]
    for k in ("tcw", "vel", "bg", "ba", "mp_pos", "Rcw"):
        step = max(np.abs(ref[k] - np.asarray(d["kf_" + k] if k != "mp_pos" else d[k]).reshape(ref[k].shape)).max(), 1e-12) \
            if k in ("tcw", "vel", "bg", "ba", "mp_pos") else 1.0
        assert np.abs(got[k] - ref[k]).max() <= max(1e-4 * step, tol), (k, np.abs(got[k] - ref[k]).max(), step)
    assert np.allclose(got["chi2"], ref["chi2"], rtol=1e-5, atol=1e-6)
    assert np.array_equal(got["depth_pos"], ref["depth_pos"])


@pytest.mark.parametrize("n_opt,n_mp,seed,perturb", [(4, 80, 2, 1.0), (6, 300, 1, 1.0), (10, 400, 3, 1.0), (5, 150, 5, 6.0)])
def test_host_run_of_the_kernel_source_matches_the_oracle(oracle, n_opt, n_mp, seed, perturb):
    from orb_slam3_b200.optimizer import lia_debug_host
    d, _ = scenes.lia_scene(n_opt, n_mp, seed=seed, perturb=perturb)
    v = oracle.make_lia_view(d)
    ref = oracle.lia_solve(v)
    got = lia_debug_host(v)
    _compare(d, ref, got)
    if perturb > 1:
        assert ref["stats"]["trials"] >= ref["stats"]["iterations"]


def test_large_window_settings_and_bad_views(oracle):
    from orb_slam3_b200._lib import OrbError
    from orb_slam3_b200.optimizer import lia_debug_host
    d, _ = scenes.lia_scene(8, 200, seed=7)
    d["lambda_init"], d["iterations"] = 1e-2, 4                   # bLarge (Optimizer.cc:2387-2392, :2509-2513)
    v = oracle.make_lia_view(d)
    ref, got = oracle.lia_solve(v), lia_debug_host(v)
    assert ref["stats"]["iterations"] <= 4
    _compare(d, ref, got)
    v.lambda_init = 0.0
    with pytest.raises(OrbError):
        lia_debug_host(v)


def test_window_without_inertial_edges(oracle):
    """Keyframes without IMU vertices and no inertial edges: the ImuCamPose parametrisation alone (the visual
    part of the window) -- the degenerate input the core must still handle."""
    from orb_slam3_b200.optimizer import lia_debug_host
    d, _ = scenes.lia_scene(5, 200, seed=11)
    for k in ("i_kf1", "i_kf2", "i_dT", "i_last"):
        d[k] = d[k][:0]
    for k in ("i_dR", "i_dV", "i_dP", "i_JRg", "i_JVg", "i_JVa", "i_JPg", "i_JPa", "i_bias", "i_C"):
        d[k] = np.asarray(d[k])[:0]
    d["kf_has_imu"] = np.zeros_like(d["kf_has_imu"])
    v = oracle.make_lia_view(d)
    ref, got = oracle.lia_solve(v), lia_debug_host(v)
    assert ref["stats"]["dim"] == 6 * 5
    _compare(d, ref, got)
    assert np.array_equal(got["vel"], d["kf_vel"]) and np.array_equal(got["bg"], d["kf_bg"])

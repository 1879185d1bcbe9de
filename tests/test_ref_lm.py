"""The LM control law of the oracle's optimisers against THE REFERENCE'S OWN OBJECT CODE (oracle/_ref/libref_lm.so).

Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.{h,cpp} is piped unmodified into the compiler (oracle/Makefile)
over stand-ins for the interfaces it is written against (g2o::Solver, g2o::SparseOptimizer: oracle/eigencompat/g2o_unit/);
oracle/ref_lm_wrap.cpp implements those interfaces by forwarding to an orc_lm_ops table -- the operations of one of the
oracle's optimisers (linearise, solve for a lambda, oplus, push / pop) -- and exports ref_lm_driver.  Each optimiser
(LocalBundleAdjustment, PoseOptimization, LocalInertialBA) runs the restated control law (orc_lm_restated, rows a21 / 8f-2
/ 8f-4b) over the SAME operations unless it is handed another driver, so with the reference's driver plugged in everything
must come out bit for bit the same: iteration and trial counts, every lambda, every chi2, which trials were accepted, the
final lambda, the per-edge chi2 (computed from the errors of the last evaluated trial, like g2o's edges hold them) and
the resulting estimates."""
import numpy as np
import pytest

from orb_slam3_b200 import scenes


@pytest.fixture(scope="module")
def driver():
    from oracle import ref as R
    if not R.lm_available():
        pytest.skip("oracle/_ref/libref_lm.so is not built and the reference tree is absent")
    return R.lm_driver()


def _same_trace(a, b):
    assert len(a) == len(b)
    assert np.array_equal(a[:, 0], b[:, 0])      # lambda of every trial
    assert np.array_equal(a[:, 1], b[:, 1])      # robust chi2 after every trial
    assert np.array_equal(a[:, 3], b[:, 3])      # accepted / rejected


def _lba_same(oracle, driver, g, max_iters, lam0):
    gv = scenes.lba_view(g)
    a = oracle.lba_solve(gv, max_iters, lam0)
    b = oracle.lba_solve(gv, max_iters, lam0, driver=driver)
    assert a["iterations"] == b["iterations"]
    for k in ("iterations", "trials", "chi2_initial", "chi2_final", "lambda_final", "stopped"):
        assert a["stats"][k] == b["stats"][k], k
    _same_trace(a["trace"], b["trace"])
    for k in ("kf_pose", "mp_pos", "chi2", "depth_pos"):
        assert np.array_equal(a[k], b[k]), k
    return a


@pytest.mark.parametrize("kf,mp,seed", [(8, 300, 1), (5, 60, 0), (20, 1500, 3)])
def test_local_ba_optimize_10_is_the_reference_driver(oracle, driver, kf, mp, seed):
    g, _ = scenes.lba_graph(kf, mp, seed=seed)
    a = _lba_same(oracle, driver, g, 10, 0.0)
    assert a["iterations"] == 10 and a["stats"]["chi2_final"] < a["stats"]["chi2_initial"]


@pytest.mark.parametrize("seed", [4, 6, 11, 15])
def test_rejected_trials_follow_the_reference_driver(oracle, driver, seed):
    """A rough start with almost no damping: trials are rejected, lambda is multiplied by ni = 2, 4, 8 ..."""
    g, _ = scenes.lba_rough_graph(seed)
    a = _lba_same(oracle, driver, g, 10, 1e-8)
    assert (a["trace"][:, 3] == 0).any() and a["stats"]["trials"] > a["iterations"]


def test_stalled_progress_stops_like_the_reference_driver(oracle, driver):
    """ORB-SLAM3's addition to g2o (optimization_algorithm_levenberg.cpp:152-160): three consecutive iterations that improve
    chi2 by less than a thousandth terminate the optimisation."""
    g, _ = scenes.lba_graph(6, 200, seed=2, outlier_frac=0.0)
    a = _lba_same(oracle, driver, g, 60, 0.0)
    assert a["iterations"] < 60


def test_rig_window_and_user_lambda(oracle, driver):
    g, _ = scenes.lba_rig_graph(8, 300, seed=1)
    _lba_same(oracle, driver, g, 10, 0.0)
    _lba_same(oracle, driver, g, 5, 10.0)      # setUserLambdaInit
    _lba_same(oracle, driver, g, 1, 0.0)


def test_hopeless_start_exhausts_max_trials_after_failure(oracle, driver):
    """Estimates thrown far off with lambda ~ 0: ten rejected trials in a row end the optimisation (qmax == 10 -> Terminate)."""
    g, _ = scenes.lba_rough_graph(7)
    rng = np.random.default_rng(0)
    free = np.nonzero(g["kf_fixed"] == 0)[0]
    g["kf_pose"][free, 4:] += rng.normal(0, 30.0, (len(free), 3))
    g["mp_pos"] += rng.normal(0, 30.0, g["mp_pos"].shape)
    a = _lba_same(oracle, driver, g, 10, 1e-30)
    assert a["iterations"] < 10 and (a["trace"][-10:, 3] == 0).all()


def test_force_stop_flag(oracle, driver):
    """pbStopFlag set before the call: optimize() returns without an iteration under either driver."""
    g, _ = scenes.lba_graph(5, 60, seed=0)
    gv = scenes.lba_view(g)
    stop = np.ones(1, np.uint8)
    a = oracle.lba_solve(gv, 10, 0.0, stop=stop)
    b = oracle.lba_solve(gv, 10, 0.0, stop=stop, driver=driver)
    assert a["iterations"] == b["iterations"] == 0 and a["stats"]["stopped"] == b["stats"]["stopped"] == 1
    assert np.array_equal(a["kf_pose"], b["kf_pose"])


@pytest.mark.parametrize("n,seed,stereo_frac,outliers", [(400, 7, 0.5, 0.1), (60, 1, 0.0, 0.3), (1500, 3, 1.0, 0.05), (9, 2, 0.5, 0.0)])
def test_pose_optimization_rounds_run_by_the_reference_driver(oracle, driver, n, seed, stereo_frac, outliers):
    """Optimizer::PoseOptimization: four rounds of optimize(10) from the frame pose, edges re-classified in between, the
    robust kernel dropped for the last round (Optimizer.cc:1000-1114) -- every round's LM run by the reference's driver."""
    v, _ = scenes.pose_scene(n, seed=seed, stereo_frac=stereo_frac, outlier_frac=outliers)
    a = oracle.pose_optimize(v)
    b = oracle.pose_optimize(v, driver=driver)
    assert a["inliers"] == b["inliers"] and np.array_equal(a["stats"], b["stats"])
    _same_trace(a["trace"], b["trace"])
    assert np.array_equal(a["pose"], b["pose"]) and np.array_equal(a["outlier"], b["outlier"]) and np.array_equal(a["chi2"], b["chi2"])
    assert a["stats"][2] >= a["stats"][1] > 0


@pytest.mark.parametrize("n_opt,n_mp,seed", [(4, 60, 2), (6, 300, 0), (8, 400, 5)])
def test_local_inertial_ba_is_run_by_the_reference_driver(oracle, driver, n_opt, n_mp, seed):
    """Optimizer::LocalInertialBA's optimize(opt_it) with setUserLambdaInit (Optimizer.cc:2503-2520)."""
    d, _ = scenes.lia_scene(n_opt, n_mp, seed=seed)
    v = oracle.make_lia_view(d)
    a = oracle.lia_solve(v)
    b = oracle.lia_solve(v, driver=driver)
    assert a["stats"] == b["stats"]
    _same_trace(a["trace"], b["trace"])
    for k in ("Rcw", "tcw", "vel", "bg", "ba", "mp_pos", "chi2", "depth_pos"):
        assert np.array_equal(a[k], b[k]), k
    assert a["stats"]["err_end"] < a["stats"]["err"]

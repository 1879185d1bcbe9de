"""The LM control law of the oracle against THE REFERENCE'S OWN OBJECT CODE (oracle/_ref/libref_lm.so).

Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.{h,cpp} is piped unmodified into the compiler (oracle/Makefile)
over stand-ins for the interfaces it is written against (g2o::Solver, g2o::SparseOptimizer: oracle/eigencompat/g2o_unit/);
oracle/ref_lm_wrap.cpp implements those interfaces by forwarding to the oracle's Stepper operations (linearise, Schur +
LDL^T solve, oplus, push / pop).  orc_lba_solve runs its restated control law (row a21, orc_lba.cpp) over the SAME
operations, so the two must agree bit for bit: iteration count, trial count, every lambda, every chi2, which trials were
accepted, the final lambda and the resulting poses and landmarks."""
import numpy as np
import pytest

from orb_slam3_b200 import scenes


@pytest.fixture(scope="module")
def ref():
    from oracle import ref as R
    if not R.lm_available():
        pytest.skip("oracle/_ref/libref_lm.so is not built and the reference tree is absent")
    return R


def _same(oracle, ref, g, max_iters, lam0):
    gv = scenes.lba_view(g)
    a = oracle.lba_solve(gv, max_iters, lam0)
    b = ref.lm_optimize(gv, max_iters, lam0)
    assert a["iterations"] == b["iterations"]
    assert a["stats"]["trials"] == b["trials"]
    assert np.array_equal(a["trace"][:, 0], b["trace"][:, 0])      # lambda of every trial
    assert np.array_equal(a["trace"][:, 1], b["trace"][:, 1])      # robust chi2 after every trial
    assert np.array_equal(a["trace"][:, 3], b["trace"][:, 3])      # accepted / rejected
    assert a["stats"]["chi2_final"] == b["chi2_final"]
    assert a["stats"]["lambda_final"] == b["lambda_final"]
    assert np.array_equal(a["kf_pose"], b["kf_pose"]) and np.array_equal(a["mp_pos"], b["mp_pos"])
    return a


@pytest.mark.parametrize("kf,mp,seed", [(8, 300, 1), (5, 60, 0), (20, 1500, 3)])
def test_optimize_10_is_the_reference_driver(oracle, ref, kf, mp, seed):
    g, _ = scenes.lba_graph(kf, mp, seed=seed)
    a = _same(oracle, ref, g, 10, 0.0)
    assert a["iterations"] == 10 and a["stats"]["chi2_final"] < a["stats"]["chi2_initial"]


@pytest.mark.parametrize("seed", [4, 6, 11, 15])
def test_rejected_trials_follow_the_reference_driver(oracle, ref, seed):
    """A rough start with almost no damping: trials are rejected, lambda is multiplied by ni = 2, 4, 8 ..."""
    g, _ = scenes.lba_rough_graph(seed)
    a = _same(oracle, ref, g, 10, 1e-8)
    assert (a["trace"][:, 3] == 0).any() and a["stats"]["trials"] > a["iterations"]


def test_stalled_progress_stops_like_the_reference_driver(oracle, ref):
    """ORB-SLAM3's addition to g2o (optimization_algorithm_levenberg.cpp:152-160): three consecutive iterations that improve
    chi2 by less than a thousandth terminate the optimisation."""
    g, _ = scenes.lba_graph(6, 200, seed=2, outlier_frac=0.0)
    a = _same(oracle, ref, g, 60, 0.0)
    assert a["iterations"] < 60


def test_rig_window_and_user_lambda(oracle, ref):
    g, _ = scenes.lba_rig_graph(8, 300, seed=1)
    _same(oracle, ref, g, 10, 0.0)
    _same(oracle, ref, g, 5, 10.0)      # setUserLambdaInit
    _same(oracle, ref, g, 1, 0.0)


def test_hopeless_start_exhausts_max_trials_after_failure(oracle, ref):
    """Poses thrown far off with lambda ~ 0: ten rejected trials in a row end the optimisation (qmax == 10 -> Terminate)."""
    g, _ = scenes.lba_rough_graph(7)
    rng = np.random.default_rng(0)
    free = np.nonzero(g["kf_fixed"] == 0)[0]
    g["kf_pose"][free, 4:] += rng.normal(0, 30.0, (len(free), 3))
    g["mp_pos"] += rng.normal(0, 30.0, g["mp_pos"].shape)
    a = _same(oracle, ref, g, 10, 1e-30)
    assert a["stats"]["trials"] >= a["iterations"]

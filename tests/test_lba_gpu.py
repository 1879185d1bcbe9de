"""GPU parity of lba_solve (C ABI) against the fp64 CPU oracle: pose / landmark
deltas within 1e-4 relative (BASELINE.json north_star), same number of LM
iterations and lambda trials, same outlier classification."""
import numpy as np
import pytest

from orb_slam3_b200 import scenes

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _compare(g, ref, got, ctx="", chi_tol=1e-6):
    assert got["iterations"] == ref["iterations"], (ctx, got["iterations"], ref["iterations"])
    assert got["stats"]["trials"] == ref["stats"]["trials"], ctx
    for key in ("kf_pose", "mp_pos"):
        d_ref = ref[key] - (g[key] if key == "mp_pos" else ref[key] * 0)
        scale = max(np.abs(ref[key] - g[key]).max(), 1e-9) if key == "mp_pos" else 1.0
        err = np.abs(got[key] - ref[key]).max()
        assert err <= TOL * scale, (ctx, key, err, scale)
    # relative error on the deltas themselves
    dref = ref["mp_pos"] - g["mp_pos"]
    dgot = got["mp_pos"] - g["mp_pos"]
    rel = np.linalg.norm(dgot - dref) / max(np.linalg.norm(dref), 1e-30)
    assert rel < TOL, (ctx, "delta points rel", rel)
    tref = ref["kf_pose"][:, 4:] - g["kf_pose"][:, 4:]
    tgot = got["kf_pose"][:, 4:] - g["kf_pose"][:, 4:]
    rel = np.linalg.norm(tgot - tref) / max(np.linalg.norm(tref), 1e-30)
    assert rel < TOL, (ctx, "delta translation rel", rel)
    assert abs(got["stats"]["chi2_final"] - ref["stats"]["chi2_final"]) <= 1e-2 * chi_tol * ref["stats"]["chi2_final"]
    assert np.allclose(got["chi2"], ref["chi2"], rtol=chi_tol, atol=chi_tol)
    assert np.array_equal(got["depth_pos"], ref["depth_pos"])


@pytest.fixture(scope="module")
def lba():
    from orb_slam3_b200.optimizer import LocalBundleAdjustment
    return LocalBundleAdjustment()


@pytest.mark.parametrize("K,L,seed", [(5, 60, 0), (10, 500, 1), (20, 3000, 2), (37, 2000, 3)])
def test_lba_matches_oracle(oracle, lba, K, L, seed):
    g, _ = scenes.lba_graph(K, L, seed=seed)
    gv = scenes.lba_view(g)
    _compare(g, oracle.lba_solve(gv), lba(gv), "K%d" % K)


def test_lba_config4_50kf_20k_landmarks(oracle, lba):
    g, _ = scenes.lba_graph(50, 20000, seed=0)
    gv = scenes.lba_view(g)
    ref, got = oracle.lba_solve(gv), lba(gv)
    _compare(g, ref, got, "config4")
    from orb_slam3_b200.optimizer import LocalBundleAdjustment as LBA
    assert np.array_equal(LBA.outliers(g, ref), LBA.outliers(g, got))


def test_lba_config5_200kf_80k_landmarks(oracle, lba):
    """BASELINE.json configs[4] at full size against the oracle (the reference's own iteration budget:
    optimize(10), Optimizer.cc:1410-1411)."""
    g, _ = scenes.lba_graph(200, 80000, seed=0)
    gv = scenes.lba_view(g)
    ref, got = oracle.lba_solve(gv), lba(gv)
    _compare(g, ref, got, "config5")
    from orb_slam3_b200.optimizer import LocalBundleAdjustment as LBA
    assert np.array_equal(LBA.outliers(g, ref), LBA.outliers(g, got))


@pytest.mark.parametrize("K,L,seed,model1,model2,mono_only", [(5, 60, 2, 1, 1, False), (12, 800, 3, 1, 1, False),
                                                                 (12, 800, 4, 0, 1, False), (12, 800, 5, 1, 0, False),
                                                                 (12, 800, 6, 0, 0, False), (12, 800, 7, 1, 0, True),
                                                                 (40, 6000, 8, 1, 1, False)])
def test_lba_second_camera_and_fisheye_edges(oracle, lba, K, L, seed, model1, model2, mono_only):
    """EdgeSE3ProjectXYZToBody edges of a two-camera rig (OptimizableTypes.cpp:192-213,
    Optimizer.cc:1366-1400) and KannalaBrandt8 cameras on the mono edges (:1326), every combination of the two camera
    models; a keyframe's left and right observation of one landmark share the pose block.  The device evaluates
    KannalaBrandt8::project's float atan2f in fp64 and rounds once, so single residuals may differ from the host
    libm's by an ulp of theta or psi (a few 1e-5 px; 2 r dr / sigma^2 is then up to ~1e-3 in an edge's chi2): chi2 is
    compared at 1e-3, the LM deltas at the 1e-4 bar as everywhere."""
    g, _ = scenes.lba_rig_graph(K, L, seed=seed, model1=model1, model2=model2, mono_only=mono_only)
    gv = scenes.lba_view(g)
    ref, got = oracle.lba_solve(gv), lba(gv)
    kb8 = model1 == 1 or (model2 == 1 and not mono_only)
    _compare(g, ref, got, "rig K%d" % K, chi_tol=1e-3 if kb8 else 1e-6)
    from orb_slam3_b200.optimizer import LocalBundleAdjustment as LBA
    margin = np.abs(ref["chi2"] - 5.991) > 1e-2   # an edge sitting on the threshold may fall either way
    assert np.array_equal(LBA.outliers(g, ref)[margin], LBA.outliers(g, got)[margin])
    if not mono_only:
        assert (g["e_stereo"] == 2).any()


def test_lba_rig_views_are_validated(lba):
    from orb_slam3_b200._lib import OrbError
    g, _ = scenes.lba_rig_graph(5, 60, seed=2)
    h = dict(g)
    del h["kf_trl"]
    with pytest.raises(OrbError):
        lba(scenes.lba_view(h))          # body edges without Trl
    h = dict(g)
    del h["kf_cam_dist"]
    with pytest.raises(OrbError):
        lba(scenes.lba_view(h))          # a KannalaBrandt8 model without its distortion coefficients
    h = dict(g)
    h["e_stereo"] = np.where(g["e_stereo"] == 2, 3, g["e_stereo"]).astype(np.uint8)
    with pytest.raises(OrbError):
        lba(scenes.lba_view(h))          # unknown edge type


def test_lba_envelope_solver_and_keyframe_order(oracle, lba):
    """The reduced solve runs inside the row envelope of S (one CTA) when the window is chain-like; a shuffled
    keyframe list is renumbered (reverse Cuthill-McKee) back to a narrow profile.  Both equal the oracle."""
    g, _ = scenes.lba_graph(50, 4000, seed=4)
    got = lba(scenes.lba_view(g))
    assert got["stats"]["solver_kind"] in (1, 2, 3) and got["stats"]["envelope_rows_max"] <= 6 * 16 + 32
    _compare(g, oracle.lba_solve(scenes.lba_view(g)), got, "natural order")
    gs = scenes.permute_keyframes(g, np.random.default_rng(3).permutation(len(g["kf_fixed"])))
    got_s = lba(scenes.lba_view(gs))
    assert got_s["stats"]["solver_kind"] in (1, 2, 3) and got_s["stats"]["envelope_rows_max"] <= 6 * 24 + 32, got_s["stats"]
    _compare(gs, oracle.lba_solve(scenes.lba_view(gs)), got_s, "shuffled keyframes")
    # a short window: every keyframe pair shares landmarks, the envelope is the whole triangle of S
    gd, _ = scenes.lba_graph(12, 800, seed=6)
    got_d = lba(scenes.lba_view(gd))
    assert got_d["stats"]["envelope_rows_max"] >= 6 * 12 - 8
    _compare(gd, oracle.lba_solve(scenes.lba_view(gd)), got_d, "dense coupling")


@pytest.mark.parametrize("mode", ["dense", "sky", "win", "win2"])
def test_lba_every_reduced_solver_kernel(mode):
    """ORB_B200_LDLT pins the reduced-solve kernel (read once per process -> subprocess): all of them must agree
    with the oracle on a chain-like window and on BASELINE.json configs[3].  win2 = the window kernel from both ends
    (two CTAs + separator); the 9-keyframe window has no separator and stays one-sided."""
    import os
    import subprocess
    import sys
    code = (
        "import numpy as np, sys\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from orb_slam3_b200 import scenes\n"
        "from orb_slam3_b200.optimizer import LocalBundleAdjustment\n"
        "from oracle import oracle\n"
        "from test_lba_gpu import _compare\n"
        "lba = LocalBundleAdjustment()\n"
        "for K, L, seed in ((9, 400, 1), (37, 2000, 3), (50, 20000, 0)):\n"
        "    g, _ = scenes.lba_graph(K, L, seed=seed)\n"
        "    got = lba(scenes.lba_view(g))\n"
        "    assert got['stats']['solver_kind'] == (%d if K > 9 or %d != 3 else 2), got['stats']\n"
        "    _compare(g, oracle.lba_solve(scenes.lba_view(g)), got, 'K%%d' %% K)\n"
        "# keyframes listed in covisibility order instead of along the trajectory (renumbered by reverse Cuthill-McKee:\n"
        "# pose blocks enter the envelope in uneven steps), and a rig window with two edges per landmark and keyframe\n"
        "g, _ = scenes.lba_graph(60, 9000, seed=5)\n"
        "g = scenes.permute_keyframes(g, np.random.default_rng(1).permutation(len(g['kf_fixed'])))\n"
        "got = lba(scenes.lba_view(g))\n"
        "_compare(g, oracle.lba_solve(scenes.lba_view(g)), got, 'shuffled')\n"
        "g, _ = scenes.lba_rig_graph(40, 6000, seed=8)\n"
        "got = lba(scenes.lba_view(g))\n"
        "_compare(g, oracle.lba_solve(scenes.lba_view(g)), got, 'rig', chi_tol=1e-3)\n"
        "print('SOLVER_OK')\n") % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)),
                                    {"dense": 0, "sky": 1, "win": 2, "win2": 3}[mode], {"dense": 0, "sky": 1, "win": 2, "win2": 3}[mode])
    env = dict(os.environ, ORB_B200_LDLT=mode)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "SOLVER_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]


def test_lba_rejected_trials_and_user_lambda(oracle, lba):
    g, _ = scenes.lba_rough_graph(4)
    gv = scenes.lba_view(g)
    ref = oracle.lba_solve(gv, max_iters=4, lambda_init=1e-8)
    got = lba(gv, max_iters=4, lambda_init=1e-8)
    assert (ref["trace"][:, 3] == 0).any()
    assert got["stats"]["trials"] == ref["stats"]["trials"] and got["iterations"] == ref["iterations"]
    assert abs(got["stats"]["chi2_final"] - ref["stats"]["chi2_final"]) <= 1e-5 * ref["stats"]["chi2_final"]
    # inertial maps: setUserLambdaInit(100) (Optimizer.cc:1197-1198)
    g2, _ = scenes.lba_graph(8, 300, seed=5)
    gv2 = scenes.lba_view(g2)
    _compare(g2, oracle.lba_solve(gv2, lambda_init=100.0), lba(gv2, lambda_init=100.0), "lambda100")


def test_lba_stop_flag_and_fixed_poses(oracle, lba):
    g, _ = scenes.lba_graph(8, 200, seed=3)
    gv = scenes.lba_view(g)
    stop = np.ones(1, np.uint8)
    r = lba(gv, pbStopFlag=stop)
    assert r["iterations"] == 0 and r["stats"]["stopped"] == 1
    assert np.allclose(r["mp_pos"], g["mp_pos"])
    r = lba(gv)
    fixed = g["kf_fixed"] == 1
    q0 = g["kf_pose"][fixed].copy()
    q0[:, :4] /= np.linalg.norm(q0[:, :4], axis=1, keepdims=True)
    assert np.allclose(r["kf_pose"][fixed], q0, atol=1e-14)
    # edge order must not matter: shuffle the edges, same result (deterministic reductions aside)
    perm = np.random.default_rng(0).permutation(len(g["e_kf"]))
    g2 = dict(g)
    for k in ("e_kf", "e_mp", "e_stereo", "e_obs", "e_inv_sigma2"):
        g2[k] = g[k][perm]
    r2 = lba(scenes.lba_view(g2))
    # (summation order inside a landmark changes -> last-bit differences, amplified by 10 LM steps)
    assert np.abs(r2["mp_pos"] - r["mp_pos"]).max() < 1e-6, np.abs(r2["mp_pos"] - r["mp_pos"]).max()
    assert np.allclose(r2["chi2"], r["chi2"][perm], rtol=1e-6, atol=1e-6)
    # same input twice: bitwise identical (all reductions are fixed-order; detects races)
    r3 = lba(gv)
    assert np.array_equal(r3["mp_pos"], r["mp_pos"]) and np.array_equal(r3["kf_pose"], r["kf_pose"])


def test_lba_landmark_shards_over_nccl_match_single_gpu():
    """Needs >= 2 GPUs: the sharded solve (one ncclAllReduce of [S | b_s] per trial) equals the 1-GPU solve."""
    import os
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(here, "multi_gpu_lba_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTI_GPU_LBA_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]

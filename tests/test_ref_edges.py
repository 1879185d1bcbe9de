"""The oracle's bundle-adjustment edges against THE REFERENCE'S OWN OBJECT CODE (oracle/_ref/libref_edges.so):
/root/reference/src/OptimizableTypes.cpp, src/CameraModels/Pinhole.cpp and KannalaBrandt8.cpp compiled unmodified
(oracle/Makefile) over a functional stand-in for the slice of Eigen / g2o they use (oracle/eigencompat/: the image has
no Eigen).  Residuals (computeError), Jacobians (linearizeOplus), isDepthPositive of
  EdgeSE3ProjectXYZ            (LocalBundleAdjustment's mono edge),
  EdgeSE3ProjectXYZToBody      (its second-camera edge, a17),
  EdgeSE3ProjectXYZOnlyPose    (PoseOptimization's mono edge, 8f-2),
and project / projectJac of both camera models are the reference's control flow and formulas as object code; what is
NOT the reference's here is the matrix arithmetic underneath (plain loops instead of Eigen's expression templates:
same operations, possibly in another order), hence 1e-9 relative instead of bit equality.
The stereo edges -- g2o::EdgeStereoSE3ProjectXYZ (a16) and g2o::EdgeStereoSE3ProjectXYZOnlyPose (8f-2) -- live in the
vendored g2o: oracle/_ref/libref_g2o.so is Thirdparty/g2o/g2o/types/types_six_dof_expmap.{h,cpp} piped unmodified into the
compiler over stand-ins for the g2o core headers they include (oracle/Makefile), held to the same 1e-9.
The libraries are built where /root/reference exists and travel prebuilt to the GPU box; skipped when absent."""
import numpy as np
import pytest

from orb_slam3_b200 import scenes

PIN = np.array([700.0, 705.0, 640.0, 360.0, 0, 0, 0, 0], np.float32)
KB8 = np.array([190.97, 190.44, 254.93, 256.89, 0.0034823894, 0.0007150348, -0.0020532361, 0.00020293673], np.float32)


@pytest.fixture(scope="module")
def ref():
    from oracle import ref as R
    if not R.edges_available():
        pytest.skip("oracle/_ref/libref_edges.so is not built and the reference tree is absent")
    return R


def _close(a, b, tol=1e-9):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max())


def test_camera_models_project_and_jacobian(ref):
    """The reference's project() / projectJac() object code against an independent numpy restatement (the one the
    finite-difference LM of test_lba_oracle.py uses) -- this also shows the stand-in arithmetic does what it says."""
    from test_lba_oracle import _project
    rng = np.random.default_rng(0)
    for model, p in ((0, PIN), (1, KB8)):
        for _ in range(200):
            X = np.array([rng.uniform(-3, 3), rng.uniform(-2, 2), rng.uniform(1.5, 30)])
            uv = ref.cam_project(model, p, X)
            assert _close(uv, _project(model, p, X, exact_float=True), 2e-6 if model else 1e-12)   # KB8: float atan2f, host libm vs numpy
            J = ref.cam_project_jac(model, p, X)
            h = 1e-6
            Jn = np.stack([(_project(model, p, X + h * np.eye(3)[c], False) - _project(model, p, X - h * np.eye(3)[c], False)) / (2 * h)
                           for c in range(3)], 1)
            assert np.abs(J - Jn).max() <= 1e-5 * max(1.0, np.abs(Jn).max())


@pytest.mark.parametrize("model1,model2", [(0, 0), (1, 1), (0, 1), (1, 0)])
def test_lba_edges_equal_the_reference_object_code(oracle, ref, model1, model2):
    """Every mono and second-camera edge of a rig window: err, d err / d point, d err / d pose, isDepthPositive."""
    g, _ = scenes.lba_rig_graph(6, 120, seed=3, model1=model1, model2=model2)
    gv = scenes.lba_view(g)
    n_mono = n_body = 0
    for e in range(0, len(g["e_kf"]), 3):
        k, l, typ = int(g["e_kf"][e]), int(g["e_mp"][e]), int(g["e_stereo"][e])
        pose = g["kf_pose"][k].copy()
        pose[:4] /= np.linalg.norm(pose[:4])
        if typ == 2:
            p8 = g["kf_cam2"][k]
            r_err, r_A, r_B, r_dp = ref.edge_binary(int(g["kf_cam2_model"][k]), p8, pose, g["mp_pos"][l], g["e_obs"][e][:2], g["kf_trl"][k])
            n_body += 1
        else:
            p8 = np.concatenate([g["kf_cam"][k][:4], g["kf_cam_dist"][k]])
            r_err, r_A, r_B, r_dp = ref.edge_binary(int(g["kf_cam_model"][k]), p8, pose, g["mp_pos"][l], g["e_obs"][e][:2])
            n_mono += 1
        err, A, B, dp = oracle.lba_edge(gv, e)
        assert _close(err[:2], r_err) and err[2] == 0, (e, typ, err, r_err)
        assert _close(A[:2], r_A), (e, typ, A[:2], r_A)
        assert _close(B[:2], r_B), (e, typ, B[:2], r_B)
        assert dp == r_dp
    assert n_mono > 50 and n_body > 30


def test_pinhole_window_edges_equal_the_reference_object_code(oracle, ref):
    """The mono edges of a plain Pinhole window (the stereo edges are g2o's: see the module docstring)."""
    g, _ = scenes.lba_graph(6, 150, seed=2, stereo_frac=0.5)
    gv = scenes.lba_view(g)
    n = 0
    for e in np.nonzero(g["e_stereo"] == 0)[0][::2]:
        k, l = int(g["e_kf"][e]), int(g["e_mp"][e])
        pose = g["kf_pose"][k].copy()
        pose[:4] /= np.linalg.norm(pose[:4])
        p8 = np.concatenate([g["kf_cam"][k][:4], np.zeros(4, np.float32)])
        r_err, r_A, r_B, r_dp = ref.edge_binary(0, p8, pose, g["mp_pos"][l], g["e_obs"][e][:2])
        err, A, B, dp = oracle.lba_edge(gv, int(e))
        assert _close(err[:2], r_err) and _close(A[:2], r_A) and _close(B[:2], r_B) and dp == r_dp, e
        n += 1
    assert n > 100


def test_pose_optimization_mono_edges_equal_the_reference_object_code(oracle, ref):
    """EdgeSE3ProjectXYZOnlyPose (Optimizer.cc:868-893) of a PoseOptimization frame."""
    v, _ = scenes.pose_scene(300, seed=4, stereo_frac=0.5)
    import ctypes as C
    xw = np.ctypeslib.as_array(C.cast(v.xw, C.POINTER(C.c_float)), (v.n * 3,)).reshape(-1, 3).astype(np.float64)
    obs = np.ctypeslib.as_array(C.cast(v.obs, C.POINTER(C.c_float)), (v.n * 3,)).reshape(-1, 3).astype(np.float64)
    pose = np.array(list(v.pose))
    pose[:4] /= np.linalg.norm(pose[:4])
    p8 = np.array([v.fx, v.fy, v.cx, v.cy, 0, 0, 0, 0], np.float32)
    n = 0
    for e in range(v.n):
        if obs[e, 2] >= 0:
            continue
        r_err, r_J, r_dp = ref.edge_unary(0, p8, pose, xw[e], obs[e, :2])
        err, B = oracle.pose_edge(v, e)
        assert _close(err[:2], r_err) and _close(B[:2], r_J), e
        n += 1
    assert n > 100


@pytest.fixture(scope="module")
def g2o(ref):
    if not ref.g2o_available():
        pytest.skip("oracle/_ref/libref_g2o.so is not built and the reference tree is absent")
    return ref


def test_lba_stereo_edges_equal_the_vendored_g2o_object_code(oracle, g2o):
    """g2o::EdgeStereoSE3ProjectXYZ (row a16: float invz in cam_project, the 3x3 / 3x6 Jacobians) on every stereo edge of a
    window, and g2o's own EdgeSE3ProjectXYZ on the mono ones (the formulas OptimizableTypes.cpp generalises)."""
    g, _ = scenes.lba_graph(6, 150, seed=5, stereo_frac=0.6)
    gv = scenes.lba_view(g)
    n_st = n_mono = 0
    for e in range(len(g["e_kf"])):
        k, l, st = int(g["e_kf"][e]), int(g["e_mp"][e]), int(g["e_stereo"][e])
        pose = g["kf_pose"][k].copy()
        pose[:4] /= np.linalg.norm(pose[:4])
        info = float(g["e_inv_sigma2"][e])
        r_err, r_A, r_B, r_dp, r_chi = g2o.g2o_edge_binary(st, g["kf_cam"][k], pose, g["mp_pos"][l], g["e_obs"][e], info)
        err, A, B, dp = oracle.lba_edge(gv, e)
        d = 3 if st else 2
        assert _close(err[:d], r_err), (e, st, err, r_err)
        assert _close(A[:d], r_A), (e, st, A[:d], r_A)
        assert _close(B[:d], r_B), (e, st, B[:d], r_B)
        assert dp == r_dp
        assert _close(info * float(err[:d] @ err[:d]), r_chi)
        if not st:
            assert err[2] == 0 and not A[2].any() and not B[2].any()
        n_st += st
        n_mono += 1 - st
    assert n_st > 300 and n_mono > 200


def test_pose_optimization_edges_equal_the_vendored_g2o_object_code(oracle, g2o):
    """g2o::EdgeStereoSE3ProjectXYZOnlyPose (Optimizer.cc:897-935) and g2o's EdgeSE3ProjectXYZOnlyPose of a frame."""
    import ctypes as C
    v, _ = scenes.pose_scene(400, seed=6, stereo_frac=0.6)
    xw = np.ctypeslib.as_array(C.cast(v.xw, C.POINTER(C.c_float)), (v.n * 3,)).reshape(-1, 3).astype(np.float64)
    obs = np.ctypeslib.as_array(C.cast(v.obs, C.POINTER(C.c_float)), (v.n * 3,)).reshape(-1, 3).astype(np.float64)
    pose = np.array(list(v.pose))
    pose[:4] /= np.linalg.norm(pose[:4])
    k5 = np.array([v.fx, v.fy, v.cx, v.cy, v.bf])
    n_st = n_mono = 0
    for e in range(v.n):
        st = int(obs[e, 2] >= 0)
        r_err, r_J, r_dp, _ = g2o.g2o_edge_unary(st, k5, pose, xw[e], obs[e])
        err, B = oracle.pose_edge(v, e)
        d = 3 if st else 2
        assert _close(err[:d], r_err), (e, st, err, r_err)
        assert _close(B[:d], r_J), (e, st)
        n_st += st
        n_mono += 1 - st
    assert n_st > 150 and n_mono > 100


def test_huber_kernel_equals_the_vendored_g2o_object_code(oracle, g2o):
    """RobustKernelHuber::setDelta / robustify (robust_kernel_impl.cpp:65-91, float dsqr): rho and rho' bit for bit, for
    both thresholds of LocalBundleAdjustment (Optimizer.cc:1275-1276), on both sides of and exactly at delta^2."""
    rng = np.random.default_rng(0)
    for th in (np.float32(np.sqrt(5.991)), np.float32(np.sqrt(7.815))):
        dsqr = float(np.float32(float(th) * float(th)))
        es = np.concatenate([rng.uniform(0, 3 * dsqr, 500), rng.uniform(0, 1e4, 200), [0.0, dsqr, np.nextafter(dsqr, 0), np.nextafter(dsqr, 1e9)]])
        for e in es:
            r = g2o.g2o_huber(float(th), float(e))
            o = oracle.huber(th, float(e))
            assert o[0] == r[0] and o[1] == r[1], (float(th), e, o, r)


def test_vertex_oplus_equals_the_vendored_g2o_object_code(oracle, g2o):
    """VertexSE3Expmap::oplusImpl = SE3Quat::exp(update) * estimate (row a22) with the reference's own se3quat.h: the
    Rodrigues / V formulas, the theta < 1e-5 branch, Quaterniond(R) in all four branches (rotations up to pi), operator*
    and normalizeRotation -- against the oracle's se3_exp_mul."""
    rng = np.random.default_rng(1)
    n_small = n_big = 0
    for i in range(3000):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        pose = np.concatenate([q if q[3] > 0 else -q, rng.normal(0, 5, 3)])
        if i % 3 == 0:      # LM-sized steps
            up = rng.normal(0, 1, 6) * 10.0 ** rng.uniform(-9, -2)
        elif i % 3 == 1:    # the small-angle branch with a large translation part
            up = np.concatenate([rng.normal(0, 1, 3) * 1e-6, rng.normal(0, 1, 3)])
        else:               # large rotations: trace(R) <= 0 picks one of the three other branches of Quaterniond(R)
            w = rng.normal(size=3)
            up = np.concatenate([w / np.linalg.norm(w) * rng.uniform(2.0, np.pi), rng.normal(0, 1, 3)])
        n_small += np.linalg.norm(up[:3]) < 1e-5
        n_big += np.linalg.norm(up[:3]) > 2
        a, b = oracle.se3_oplus(pose, up), g2o.g2o_oplus(pose, up)
        assert np.abs(a - b).max() <= 1e-12 * max(1.0, np.abs(b).max()), (i, up, a, b)
        assert abs(np.linalg.norm(b[:4]) - 1) < 1e-12 and b[3] >= 0
    assert n_small > 500 and n_big > 500


@pytest.mark.parametrize("kf,mp,seed,stereo_frac", [(8, 300, 1, 0.8), (6, 150, 5, 0.0), (12, 600, 2, 1.0)])
def test_normal_equations_equal_g2o_construct_quadratic_form(oracle, g2o, kf, mp, seed, stereo_frac):
    """Row a18: every edge's contribution to the Hessian blocks and the right-hand side -- information, the Huber weight
    rho'(chi2) on both, the off-diagonal block A' wOmega B -- accumulated edge by edge by BaseBinaryEdge::
    constructQuadraticForm (base_binary_edge.hpp:57-120) as object code, against the oracle's build_system()."""
    g, _ = scenes.lba_graph(kf, mp, seed=seed, stereo_frac=stereo_frac)   # 2 % gross outliers: the kernel is active
    gv = scenes.lba_view(g)
    a, b = oracle.lba_system(gv), g2o.g2o_build_system(gv)
    for k in ("Hpp", "Hll", "W", "bp", "bl"):
        assert np.abs(a[k]).max() > 0
        assert np.abs(a[k] - b[k]).max() <= 1e-12 * np.abs(b[k]).max(), k
    fixed = g["kf_fixed"] != 0
    assert fixed.any() and not a["Hpp"][fixed].any() and not b["Hpp"][fixed].any()   # a fixed vertex gets no block


@pytest.mark.parametrize("n,seed,stereo_frac", [(400, 6, 0.6), (50, 2, 0.0), (900, 3, 1.0)])
def test_pose_system_equals_g2o_construct_quadratic_form(oracle, g2o, n, seed, stereo_frac):
    """8f-2: BaseUnaryEdge::constructQuadraticForm (base_unary_edge.hpp:42-72) over the frame's edges."""
    v, _ = scenes.pose_scene(n, seed=seed, stereo_frac=stereo_frac)
    H, b = oracle.pose_system(v)
    Hr, br = g2o.g2o_pose_system(v)
    assert np.abs(H - Hr).max() <= 1e-12 * np.abs(Hr).max() and np.abs(b - br).max() <= 1e-12 * np.abs(br).max()
    assert np.abs(H - H.T).max() <= 1e-12 * np.abs(H).max()


@pytest.mark.parametrize("n_opt,n_mp,seed", [(6, 300, 0), (4, 120, 3)])
def test_inertial_ba_visual_edges_equal_the_reference_object_code(oracle, n_opt, n_mp, seed):
    """8f-4b: EdgeMono / EdgeStereo of LocalInertialBA (Optimizer.cc:2636-2735) -- src/G2oTypes.cc compiled unmodified into
    oracle/_ref/libref_vi.so, run on an ImuCamPose filled from the view -- computeError, linearizeOplus (the body-frame
    SE3 derivative through Rcb / tbc), isDepthPositive, on every visual edge of a window."""
    from oracle import ref as R
    if not R.vi_available():
        pytest.skip("oracle/_ref/libref_vi.so is not built and the reference tree is absent")
    d, _ = scenes.lia_scene(n_opt, n_mp, seed=seed)
    v = oracle.make_lia_view(d)
    n_st = n_mono = 0
    for e in range(v.n_edges):
        err, A, B, dp = oracle.lia_edge(v, e)
        r_err, r_A, r_B, r_dp = R.vi_edge(v, e)
        # the residual is a difference of ~1e3 px quantities: 1e-12 px absolute
        assert np.abs(err - r_err).max() <= 1e-12 * max(1.0, np.abs(d["e_obs"][e]).max()), (e, err, r_err)
        assert _close(A, r_A, 1e-12) and _close(B, r_B, 1e-12), e
        assert dp == r_dp
        st = int(d["e_stereo"][e])
        if not st:
            assert err[2] == 0 and not A[2].any() and not B[2].any()
        n_st += st
        n_mono += 1 - st
    assert n_st > 100 and n_mono > 50

"""ctypes binding of liborb_oracle.so (TEST INFRASTRUCTURE ONLY)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liborb_oracle.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28


def build(force=False):
    """Compile the oracle with the committed Makefile (gcc only)."""
    srcs = [f for f in os.listdir(_HERE) if f.startswith("orc_") or f.endswith(".inc")]
    if not force and os.path.exists(_LIB_PATH):
        so_m = os.path.getmtime(_LIB_PATH)
        deps = [os.path.join(_HERE, f) for f in srcs] + [os.path.join(_HERE, "..", "include", "orb_b200.h")]
        if all(os.path.getmtime(f) <= so_m for f in deps):
            return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "liborb_oracle.so"],
                          stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        u8p = C.POINTER(C.c_uint8)
        ip = C.POINTER(C.c_int)
        L.orc_extractor_create.restype = C.c_void_p
        L.orc_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orc_extractor_destroy.argtypes = [C.c_void_p]
        L.orc_extract.restype = C.c_int
        L.orc_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_int, ip]
        L.orc_level_info.argtypes = [C.c_void_p, C.c_int, ip, ip, ip, C.POINTER(C.c_float)]
        L.orc_level_ptr.restype = u8p
        L.orc_level_ptr.argtypes = [C.c_void_p, C.c_int]
        L.orc_level_candidates.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_level_keypoints.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_umax.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_resize_linear_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                           C.c_int, C.c_int]
        L.orc_fast.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orc_gaussian_blur7.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orc_fast_atan2.restype = C.c_float
        L.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
        L.orc_cos_sin_deg.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.orc_pattern.argtypes = [C.c_void_p]
        L.orc_distribute.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orc_extract_throughput.restype = C.c_double
        L.orc_extract_throughput.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                             C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong)]
        L.orc_ham_distance.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_match_project_local.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_float,
                                              C.c_void_p]
        L.orc_match_project_last.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float,
                                             C.c_int, C.c_void_p]
        L.orc_match_triangulate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orc_lba_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_lba_reduced_system.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.POINTER(C.c_double)]
        L.orc_sort_nodes.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_lia_solve.restype = C.c_int
        L.orc_lia_solve.argtypes = [C.c_void_p] * 6
        L.orc_lia_linearize.restype = C.c_int
        L.orc_lia_linearize.argtypes = [C.c_void_p] * 5
        L.orc_bow_transform.restype = C.c_int
        L.orc_bow_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 7 + [C.c_int]
        L.orc_is_in_frustum.restype = C.c_int
        L.orc_is_in_frustum.argtypes = [C.c_void_p, C.c_float] + [C.c_void_p] * 7
        L.orc_pose_optimize.restype = C.c_int
        L.orc_pose_optimize.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_stereo_match.restype = C.c_int
        L.orc_stereo_match.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleExtractor:
    """CPU restatement of ORB_SLAM3::ORBextractor (src/ORBextractor.cc)."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.nlevels = nlevels
        self.nfeatures = nfeatures
        self._h = lib().orc_extractor_create(nfeatures, scale_factor, nlevels, ini_th, min_th)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_extractor_destroy(self._h)
            self._h = None

    def extract(self, img, lap=(0, 0)):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        cap = self.nfeatures * 2 + 64 * self.nlevels
        kps = np.zeros(cap, dtype=KP_DTYPE)
        desc = np.zeros((cap, 32), dtype=np.uint8)
        n = C.c_int(0)
        mono = lib().orc_extract(self._h, _ptr(img), img.shape[0], img.shape[1], img.strides[0],
                                 int(lap[0]), int(lap[1]), _ptr(kps), _ptr(desc), cap, C.byref(n))
        if mono == -3:
            raise ValueError("aspect ratio < 0.5 at some pyramid level: undefined in the reference (nIni == 0, ORBextractor.cc:560)")
        if mono < 0:
            raise RuntimeError("oracle extract rc=%d" % mono)
        return kps[:n.value].copy(), desc[:n.value].copy(), mono

    def level_info(self, level):
        w, h, q = C.c_int(), C.c_int(), C.c_int()
        s = C.c_float()
        lib().orc_level_info(self._h, level, C.byref(w), C.byref(h), C.byref(q), C.byref(s))
        return w.value, h.value, q.value, s.value

    def level_image(self, level):
        w, h, _, _ = self.level_info(level)
        p = lib().orc_level_ptr(self._h, level)
        return np.ctypeslib.as_array(p, shape=(h, w)).copy()

    def level_candidates(self, level):
        n = lib().orc_level_candidates(self._h, level, None, 0)
        out = np.zeros(n, dtype=KP_DTYPE)
        lib().orc_level_candidates(self._h, level, _ptr(out), n)
        return out

    def level_keypoints(self, level):
        n = lib().orc_level_keypoints(self._h, level, None, 0)
        out = np.zeros(n, dtype=KP_DTYPE)
        lib().orc_level_keypoints(self._h, level, _ptr(out), n)
        return out

    def umax(self):
        out = np.zeros(16, dtype=np.int32)
        lib().orc_umax(self._h, _ptr(out))
        return out


def resize_linear_u8(src, dw, dh):
    src = np.ascontiguousarray(src, dtype=np.uint8)
    dst = np.zeros((dh, dw), dtype=np.uint8)
    lib().orc_resize_linear_u8(_ptr(src), src.shape[1], src.shape[0], src.strides[0], _ptr(dst), dw, dh, dw)
    return dst


def fast(img, threshold, nonmax=True):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    cap = img.size
    out = np.zeros((cap, 3), dtype=np.int32)
    n = lib().orc_fast(_ptr(img), img.shape[1], img.shape[0], img.strides[0], threshold, int(nonmax),
                       _ptr(out), cap)
    return out[:n].copy()


def gaussian_blur7(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    dst = np.zeros_like(img)
    lib().orc_gaussian_blur7(_ptr(img), img.shape[1], img.shape[0], img.strides[0], _ptr(dst), dst.strides[0])
    return dst


def fast_atan2(y, x):
    return lib().orc_fast_atan2(float(y), float(x))


def cos_sin_deg(angle):
    c, s = C.c_float(), C.c_float()
    lib().orc_cos_sin_deg(float(angle), C.byref(c), C.byref(s))
    return c.value, s.value


def pattern():
    out = np.zeros(1024, dtype=np.int32)
    lib().orc_pattern(_ptr(out))
    return out


def sort_nodes(count, ulx):
    count = np.ascontiguousarray(count, dtype=np.int32)
    ulx = np.ascontiguousarray(ulx, dtype=np.int32)
    perm = np.zeros(len(count), dtype=np.int32)
    lib().orc_sort_nodes(_ptr(count), _ptr(ulx), len(count), _ptr(perm))
    return perm


def distribute(xys, band_w, band_h, n_features):
    """DistributeOctTree on candidates xys (n,3) int32 [x, y, response]."""
    xys = np.ascontiguousarray(xys, dtype=np.int32).reshape(-1, 3)
    cap = n_features + 64
    out = np.zeros((cap, 3), dtype=np.int32)
    m = lib().orc_distribute(_ptr(xys), len(xys), band_w, band_h, n_features, _ptr(out), cap)
    return out[:m].copy()


def extract_throughput(frames, nfeatures, nthreads, iters, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
    """frames [n,H,W] uint8 -> (frames_per_second, frames_done, seconds) using C++ std::threads."""
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    tot = C.c_longlong(0)
    dt = lib().orc_extract_throughput(nfeatures, scale_factor, nlevels, ini_th, min_th, _ptr(frames),
                                      frames.shape[0], frames.shape[1], frames.shape[2], nthreads, iters,
                                      C.byref(tot))
    done = nthreads * iters
    return done / dt, done, dt


# ---- matchers: the views are the ctypes structs of orb_slam3_b200/views.py (interface types)
def three_maxima(sizes, init=(-1, -1, -1)):
    sizes = np.ascontiguousarray(sizes, np.int32)
    ind = np.array(init, np.int32)
    lib().orc_three_maxima.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib().orc_three_maxima(_ptr(sizes), len(sizes), _ptr(ind))
    return tuple(int(v) for v in ind)


def ham_distance(a, b):
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    return lib().orc_ham_distance(_ptr(a), _ptr(b))


def match_project_local(F, mps, th, nn_ratio, far_points=False, th_far=50.0):
    out = np.empty(F.n, np.int32)
    n = lib().orc_match_project_local(C.byref(F), C.byref(mps), th, nn_ratio, int(far_points), th_far, _ptr(out))
    return n, out


def match_project_last(cur, last, Tcw_qt7, th, forward=False, backward=False, check_ori=True):
    T = np.ascontiguousarray(Tcw_qt7, np.float32)
    out = np.empty(cur.n, np.int32)
    n = lib().orc_match_project_last(C.byref(cur), C.byref(last), _ptr(T), int(forward), int(backward), th,
                                     int(check_ori), _ptr(out))
    return n, out


def match_triangulate(kf1, kf2, fv1, fv2, F12, ep, only_stereo=False, coarse=False, check_ori=True, cap=None):
    cap = cap or max(kf1.n, 1)
    F12 = np.ascontiguousarray(F12, np.float32).reshape(9)
    ep = np.ascontiguousarray(ep, np.float32).reshape(2)
    out = np.empty((cap, 2), np.int32)
    n = lib().orc_match_triangulate(C.byref(kf1), C.byref(kf2), C.byref(fv1), C.byref(fv2), _ptr(F12), _ptr(ep),
                                    int(only_stereo), int(coarse), int(check_ori), _ptr(out), cap)
    return n, out[:n]


# ---- local BA (views: orb_slam3_b200/views.py lba_graph_view / lba_stats, interface types)
def lba_solve(g, max_iters=10, lambda_init=0.0, stop=None, driver=None):
    """optimize(max_iters) on the flat graph view g.  Returns dict(kf_pose, mp_pos, chi2, depth_pos, stats, trace).
    driver: address of an orc_lm_driver that runs the LM control law (None = the restated one; oracle.ref.lm_driver() =
    the reference's optimization_algorithm_levenberg.cpp as object code)."""
    from orb_slam3_b200.views import lba_stats
    kf = np.zeros((g.n_kf, 7))
    mp = np.zeros((g.n_mp, 3))
    chi2 = np.zeros(g.n_edges)
    dp = np.zeros(g.n_edges, np.uint8)
    st = lba_stats()
    trace = np.full((128, 4), np.nan)
    sp = _ptr(stop) if stop is not None else None
    it = lib().orc_lba_solve_lm(C.byref(g), sp, max_iters, C.c_double(lambda_init), _ptr(kf), _ptr(mp), _ptr(chi2), _ptr(dp),
                                C.byref(st), _ptr(trace), C.c_void_p(driver))
    return dict(iterations=it, kf_pose=kf, mp_pos=mp, chi2=chi2, depth_pos=dp, stats=st.as_dict(),
                trace=trace[:st.trials].copy())


def lba_system(g):
    """The normal equations build_system() forms at the input estimates: dict(Hpp[n_kf,6,6], Hll[n_mp,3,3], W[n_edges,6,3],
    bp[n_kf,6], bl[n_mp,3]); rows of fixed keyframes are zero."""
    out = dict(Hpp=np.zeros((g.n_kf, 6, 6)), Hll=np.zeros((g.n_mp, 3, 3)), W=np.zeros((g.n_edges, 6, 3)),
               bp=np.zeros((g.n_kf, 6)), bl=np.zeros((g.n_mp, 3)))
    lib().orc_lba_system(C.byref(g), *[_ptr(out[k]) for k in ("Hpp", "Hll", "W", "bp", "bl")])
    return out


def lba_edge(g, e):
    """One edge of an lba_graph_view at its input estimates: err[3], A[d x 3] = d err / d point, B[d x 6] = d err / d pose
    (3 rows, the third zero for 2-D edges), isDepthPositive."""
    err, A, B, dp = np.zeros(3), np.zeros((3, 3)), np.zeros((3, 6)), C.c_uint8(0)
    rc = lib().orc_lba_edge(C.byref(g), int(e), _ptr(err), _ptr(A), _ptr(B), C.byref(dp))
    assert rc == 0
    return err, A, B, bool(dp.value)


def huber(th, e):
    """RobustKernelHuber with setDelta((float)th): (rho(e), rho'(e)) of the squared error e."""
    out = np.zeros(2)
    lib().orc_huber(C.c_float(th), C.c_double(e), _ptr(out))
    return out


def se3_oplus(pose7, update6):
    """VertexSE3Expmap::oplusImpl: SE3Quat::exp(update6) * pose7 (quaternion xyzw + t), normalised like g2o does."""
    out = np.zeros(7)
    lib().orc_se3_oplus(_ptr(np.ascontiguousarray(pose7, np.float64)), _ptr(np.ascontiguousarray(update6, np.float64)), _ptr(out))
    return out


def pose_edge(v, e):
    """One edge of a pose_opt_view at its input pose: err[3], B[3 x 6] = d err / d pose."""
    err, B = np.zeros(3), np.zeros((3, 6))
    rc = lib().orc_pose_edge(C.byref(v), int(e), _ptr(err), _ptr(B))
    assert rc == 0
    return err, B


def lba_reduced_system(g, lam, lm_mask=None):
    nf = int((np.ctypeslib.as_array(C.cast(g.kf_fixed, C.POINTER(C.c_uint8)), (g.n_kf,)) == 0).sum())
    n = 6 * nf
    S = np.zeros((n, n))
    bs = np.zeros(n)
    chi = C.c_double()
    m = None if lm_mask is None else np.ascontiguousarray(lm_mask, np.uint8)
    lib().orc_lba_reduced_system(C.byref(g), lam, _ptr(m) if m is not None else None, _ptr(S), _ptr(bs),
                                 C.byref(chi))
    return S, bs, chi.value


def stereo_match(kl, dl, kr, dr, pyr_l, pyr_r, bf, b, scale_factor=1.2):
    """Frame::ComputeStereoMatches (Frame.cc:811-981).  pyr_l / pyr_r: lists of un-blurred level images
    (mvImagePyramid of the two extractors).  Returns (n_kept, mvuRight, mvDepth, sad)."""
    nl = len(pyr_l)
    L = [np.ascontiguousarray(a, np.uint8) for a in pyr_l]
    R = [np.ascontiguousarray(a, np.uint8) for a in pyr_r]
    assert all(a.shape == c.shape for a, c in zip(L, R))
    pl = (C.c_void_p * nl)(*[a.ctypes.data for a in L])
    pr = (C.c_void_p * nl)(*[a.ctypes.data for a in R])
    lw = np.array([a.shape[1] for a in L], np.int32)
    lh = np.array([a.shape[0] for a in L], np.int32)
    ls = np.array([a.strides[0] for a in L], np.int32)
    scale = np.ones(nl, np.float32)
    for i in range(1, nl):  # ORBextractor.cc:414-420
        scale[i] = np.float32(scale[i - 1] * np.float32(scale_factor))
    inv = (np.float32(1.0) / scale).astype(np.float32)
    kl = np.ascontiguousarray(kl); kr = np.ascontiguousarray(kr)
    dl = np.ascontiguousarray(dl, np.uint8); dr = np.ascontiguousarray(dr, np.uint8)
    ur = np.zeros(len(kl), np.float32)
    dp = np.zeros(len(kl), np.float32)
    sad = np.zeros(len(kl), np.int32)
    n = lib().orc_stereo_match(len(kl), _ptr(kl), _ptr(dl), len(kr), _ptr(kr), _ptr(dr), nl, pl, pr, _ptr(lw), _ptr(lh),
                               _ptr(ls), _ptr(scale), _ptr(inv), float(bf), float(b), _ptr(ur), _ptr(dp), _ptr(sad))
    return n, ur, dp, sad


def pose_system(view):
    """(H[6,6], b[6]) of PoseOptimization's first linearisation (all edges active, robust)."""
    H, b = np.zeros((6, 6)), np.zeros(6)
    lib().orc_pose_system(C.byref(view), _ptr(H), _ptr(b))
    return H, b


def pose_optimize(view, driver=None):
    """Optimizer::PoseOptimization (Optimizer.cc:814-1115) on a pose_opt_view.
    Returns dict(inliers, pose[7], outlier[n] bool, chi2[n], stats = rounds / LM iterations / LM trials, trace).
    driver: as for lba_solve."""
    pose = np.zeros(7)
    out = np.zeros(max(view.n, 1), np.uint8)
    chi2 = np.zeros(max(view.n, 1))
    stats = np.zeros(3, np.int32)
    trace = np.full((128, 4), np.nan)
    n = lib().orc_pose_optimize_lm(C.byref(view), _ptr(pose), _ptr(out), _ptr(chi2), _ptr(stats), C.c_void_p(driver), _ptr(trace))
    return dict(inliers=n, pose=pose, outlier=out[:view.n].astype(bool), chi2=chi2[:view.n], stats=stats,
                trace=trace[:min(int(stats[2]), 128)].copy())


def is_in_frustum(view, viewing_cos_limit=0.5, out=None):
    """Frame::isInFrustum (Frame.cc:512-570) over all points of an orb_frustum_view.  Returns (n_in_view, out);
    out keeps stale members for points that are not in view, like the reference."""
    if out is None:
        n = view.n
        out = dict(track_in_view=np.zeros(n, np.uint8), proj_x=np.zeros(n, np.float32), proj_y=np.zeros(n, np.float32),
                   proj_xr=np.zeros(n, np.float32), scale_level=np.zeros(n, np.int32),
                   view_cos=np.zeros(n, np.float32), depth=np.zeros(n, np.float32))
    k = lib().orc_is_in_frustum(C.byref(view), float(viewing_cos_limit),
                                *[_ptr(out[f]) for f in ("track_in_view", "proj_x", "proj_y", "proj_xr",
                                                         "scale_level", "view_cos", "depth")])
    return k, out


def bow_transform(vocab_view, desc, levelsup=4):
    """DBoW2 transform(features, BowVector, FeatureVector, levelsup) (TemplatedVocabulary.h:1127-1195).
    Returns dict(used, bow_ids, bow_vals, fv_node_ids, fv_ptr, fv_idx)."""
    desc = np.ascontiguousarray(desc, np.uint8)
    n = len(desc)
    cap = max(n, 1)
    ids, vals = np.zeros(cap, np.int32), np.zeros(cap)
    fn, fp, fi = np.zeros(cap, np.int32), np.zeros(cap + 1, np.int32), np.zeros(cap, np.int32)
    nw, nn = C.c_int(0), C.c_int(0)
    used = lib().orc_bow_transform(C.byref(vocab_view), _ptr(desc), n, int(levelsup), _ptr(ids), _ptr(vals),
                                   C.byref(nw), _ptr(fn), _ptr(fp), _ptr(fi), C.byref(nn), cap)
    assert used >= 0
    return dict(used=used, bow_ids=ids[:nw.value].copy(), bow_vals=vals[:nw.value].copy(),
                fv_node_ids=fn[:nn.value].copy(), fv_ptr=fp[:nn.value + 1].copy(), fv_idx=fi[:used].copy())


# ---- Optimizer::LocalInertialBA (oracle only so far; the view mirrors oracle/orc_lia.h)
from orb_slam3_b200.views import lia_graph_view, make_lia_view  # noqa: E402,F401  (interface types only)


def lia_solve(v, driver=None):
    """Optimizer::LocalInertialBA's optimize() on a lia_graph_view (driver: as for lba_solve)."""
    kf = np.zeros((v.n_kf, 21))
    mp = np.zeros((v.n_mp, 3))
    chi2 = np.zeros(max(v.n_edges, 1))
    dp = np.zeros(max(v.n_edges, 1), np.uint8)
    st = np.zeros(6)
    trace = np.full((128, 4), np.nan)
    it = lib().orc_lia_solve_lm(C.byref(v), _ptr(kf), _ptr(mp), _ptr(chi2), _ptr(dp), _ptr(st), C.c_void_p(driver), _ptr(trace))
    return dict(iterations=it, Rcw=kf[:, :9].reshape(-1, 3, 3), tcw=kf[:, 9:12], vel=kf[:, 12:15], bg=kf[:, 15:18],
                ba=kf[:, 18:21], mp_pos=mp, chi2=chi2[:v.n_edges], depth_pos=dp[:v.n_edges],
                stats=dict(iterations=int(st[0]), trials=int(st[1]), err=st[2], err_end=st[3], lambda_final=st[4],
                           dim=int(st[5])), trace=trace[:min(int(st[1]), 128)].copy())


def lia_edge(v, e):
    """One visual edge of a lia_graph_view at its input state: err[3], A[3,3], B[3,6], isDepthPositive."""
    err, A, B, dp = np.zeros(3), np.zeros((3, 3)), np.zeros((3, 6)), C.c_uint8(0)
    rc = lib().orc_lia_edge(C.byref(v), int(e), _ptr(err), _ptr(A), _ptr(B), C.byref(dp))
    assert rc == 0
    return err, A, B, bool(dp.value)


def lia_linearize(v, delta=None):
    """(robust chi2, b) at the input state, and robust chi2 after oplus(delta)."""
    chi, chid = C.c_double(), C.c_double()
    np_ = lib().orc_lia_linearize(C.byref(v), None, C.byref(chi), None, None)
    b = np.zeros(np_ + 3 * v.n_mp)
    d = None if delta is None else np.ascontiguousarray(delta, np.float64)
    lib().orc_lia_linearize(C.byref(v), _ptr(d) if d is not None else None, C.byref(chi), _ptr(b),
                            C.byref(chid) if d is not None else None)
    return chi.value, b, (chid.value if d is not None else None), np_

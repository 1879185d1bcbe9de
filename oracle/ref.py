"""ctypes binding of oracle/_ref/libref_orb.so (TEST INFRASTRUCTURE ONLY): the REFERENCE's own object code --
/root/reference/src/ORBextractor.cc compiled unmodified against oracle/cvcompat/ plus ORBmatcher's
DescriptorDistance / ComputeThreeMaxima -- used to pin the restated oracle (and through it the CUDA path) to
the reference itself.  Built by `make -C oracle ref` where /root/reference exists; shipped prebuilt to the GPU box."""
import ctypes as C
import os
import subprocess

import numpy as np

from . import oracle as _o

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref_orb.so")
REFERENCE = os.environ.get("ORB_REFERENCE_DIR", "/root/reference")
KP_DTYPE = _o.KP_DTYPE


def build(force=False):
    """Compile oracle/_ref from the reference sources where they lie (needs /root/reference)."""
    if not os.path.exists(os.path.join(REFERENCE, "src", "ORBextractor.cc")):
        return None
    _o.build()
    cmd = ["make", "-C", _HERE, "REF=" + REFERENCE] + (["-B"] if force else []) + ["_ref/libref_orb.so"]
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    return LIB_PATH


def available():
    return os.path.exists(LIB_PATH) or build() is not None


# ---- oracle/_ref/libref_edges.so: the reference's BA edges / camera models (OptimizableTypes.cpp, Pinhole.cpp,
#      KannalaBrandt8.cpp compiled unmodified against oracle/eigencompat/)
EDGES_LIB_PATH = os.path.join(_HERE, "_ref", "libref_edges.so")
_edges = None


def build_edges(force=False):
    if not os.path.exists(os.path.join(REFERENCE, "src", "OptimizableTypes.cpp")):
        return None
    cmd = ["make", "-C", _HERE, "REF=" + REFERENCE] + (["-B"] if force else []) + ["_ref/libref_edges.so"]
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    return EDGES_LIB_PATH


def edges_available():
    return os.path.exists(EDGES_LIB_PATH) or build_edges() is not None


def edges_lib():
    global _edges
    if _edges is None:
        if not os.path.exists(EDGES_LIB_PATH) and build_edges() is None:
            raise FileNotFoundError("oracle/_ref/libref_edges.so is not built and %s is absent" % REFERENCE)
        L = C.CDLL(EDGES_LIB_PATH)
        vp = C.c_void_p
        L.ref_cam_project.argtypes = [C.c_int, vp, vp, vp]
        L.ref_cam_project_jac.argtypes = [C.c_int, vp, vp, vp]
        L.ref_edge_binary.argtypes = [C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, C.POINTER(C.c_int)]
        L.ref_edge_unary.argtypes = [C.c_int, vp, vp, vp, vp, vp, vp, vp, C.POINTER(C.c_int)]
        _edges = L
    return _edges


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def edge_binary(model, p8, pose7, X, obs, trl7=None):
    """The reference's EdgeSE3ProjectXYZ (trl7 None) / EdgeSE3ProjectXYZToBody: computeError, linearizeOplus,
    isDepthPositive.  Returns err[2], Jxi[2,3] (point), Jxj[2,6] (pose), depth_positive."""
    p8, pose7, X, obs = (np.ascontiguousarray(p8, np.float32), np.ascontiguousarray(pose7, np.float64),
                         np.ascontiguousarray(X, np.float64), np.ascontiguousarray(obs, np.float64))
    t = None if trl7 is None else np.ascontiguousarray(trl7, np.float64)
    err, Jxi, Jxj, dp = np.zeros(2), np.zeros((2, 3)), np.zeros((2, 6)), C.c_int(0)
    edges_lib().ref_edge_binary(int(model), _p(p8), _p(pose7), None if t is None else _p(t), _p(X), _p(obs), _p(err), _p(Jxi),
                                _p(Jxj), C.byref(dp))
    return err, Jxi, Jxj, bool(dp.value)


def edge_unary(model, p8, pose7, Xw, obs, trl7=None):
    """EdgeSE3ProjectXYZOnlyPose / OnlyPoseToBody: err[2], Jxi[2,6], depth_positive."""
    p8, pose7, Xw, obs = (np.ascontiguousarray(p8, np.float32), np.ascontiguousarray(pose7, np.float64),
                          np.ascontiguousarray(Xw, np.float64), np.ascontiguousarray(obs, np.float64))
    t = None if trl7 is None else np.ascontiguousarray(trl7, np.float64)
    err, J, dp = np.zeros(2), np.zeros((2, 6)), C.c_int(0)
    edges_lib().ref_edge_unary(int(model), _p(p8), _p(pose7), None if t is None else _p(t), _p(Xw), _p(obs), _p(err), _p(J),
                               C.byref(dp))
    return err, J, bool(dp.value)


def cam_project(model, p8, X):
    p8, X, uv = np.ascontiguousarray(p8, np.float32), np.ascontiguousarray(X, np.float64), np.zeros(2)
    edges_lib().ref_cam_project(int(model), _p(p8), _p(X), _p(uv))
    return uv


def cam_project_jac(model, p8, X):
    p8, X, J = np.ascontiguousarray(p8, np.float32), np.ascontiguousarray(X, np.float64), np.zeros((2, 3))
    edges_lib().ref_cam_project_jac(int(model), _p(p8), _p(X), _p(J))
    return J


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH) and build() is None:
            raise FileNotFoundError("oracle/_ref is not built and %s is absent" % REFERENCE)
        _o.lib()  # liborb_oracle.so first: _ref's image primitives resolve against it
        L = C.CDLL(LIB_PATH)
        ip, fp = C.POINTER(C.c_int), C.POINTER(C.c_float)
        L.ref_extractor_create.restype = C.c_void_p
        L.ref_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.ref_extractor_destroy.argtypes = [C.c_void_p]
        L.ref_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                  C.c_void_p, C.c_int, ip]
        L.ref_level_info.argtypes = [C.c_void_p, C.c_int, ip, ip, ip, ip, fp]
        L.ref_level_ptr.restype = C.POINTER(C.c_uint8)
        L.ref_level_ptr.argtypes = [C.c_void_p, C.c_int]
        L.ref_tables.argtypes = [C.c_void_p] * 8
        L.ref_descriptor_distance.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_three_maxima.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ref_extract_throughput.restype = C.c_double
        L.ref_extract_throughput.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                             C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong)]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class RefExtractor:
    """ORB_SLAM3::ORBextractor itself (reference object code)."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.nlevels, self.nfeatures = nlevels, nfeatures
        self._h = lib().ref_extractor_create(nfeatures, scale_factor, nlevels, ini_th, min_th)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ref_extractor_destroy(self._h)
            self._h = None

    def extract(self, img, lap=(0, 0)):
        """Returns (keypoints, descriptors, monoIndex) of operator() (ORBextractor.cc:1086-1168)."""
        img = np.ascontiguousarray(img, dtype=np.uint8)
        cap = self.nfeatures * 2 + 64 * self.nlevels
        kps = np.zeros(cap, dtype=KP_DTYPE)
        desc = np.zeros((cap, 32), dtype=np.uint8)
        n = C.c_int(0)
        mono = lib().ref_extract(self._h, _ptr(img), img.shape[0], img.shape[1], img.strides[0], int(lap[0]), int(lap[1]),
                                 _ptr(kps), _ptr(desc), cap, C.byref(n))
        assert n.value <= cap
        return kps[:n.value].copy(), desc[:n.value].copy(), mono

    def level_image(self, level, border=0):
        """mvImagePyramid[level]; border = 19 also returns the EDGE_THRESHOLD frame around it."""
        w, h, step, q = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        s = C.c_float()
        assert lib().ref_level_info(self._h, level, C.byref(w), C.byref(h), C.byref(step), C.byref(q), C.byref(s)) == 0
        p = lib().ref_level_ptr(self._h, level)
        base = C.cast(p, C.c_void_p).value - border * step.value - border
        buf = (C.c_uint8 * ((h.value + 2 * border) * step.value)).from_address(base)
        a = np.frombuffer(buf, np.uint8).reshape(h.value + 2 * border, step.value)
        return a[:, :w.value + 2 * border].copy()

    def tables(self):
        nl = self.nlevels
        sc, isc, s2, is2 = (np.zeros(nl, np.float32) for _ in range(4))
        q, um, pat = np.zeros(nl, np.int32), np.zeros(16, np.int32), np.zeros(1024, np.int32)
        lib().ref_tables(self._h, _ptr(sc), _ptr(isc), _ptr(s2), _ptr(is2), _ptr(q), _ptr(um), _ptr(pat))
        return dict(scale=sc, inv_scale=isc, sigma2=s2, inv_sigma2=is2, quota=q, umax=um, pattern=pat)


def descriptor_distance(a, b):
    a, b = np.ascontiguousarray(a, np.uint8), np.ascontiguousarray(b, np.uint8)
    return lib().ref_descriptor_distance(_ptr(a), _ptr(b))


def three_maxima(sizes, init=(-1, -1, -1)):
    sizes = np.ascontiguousarray(sizes, np.int32)
    ind = np.array(init, np.int32)
    lib().ref_three_maxima(_ptr(sizes), len(sizes), _ptr(ind))
    return tuple(int(v) for v in ind)


def extract_throughput(frames, nfeatures, nthreads, iters):
    """Seconds for nthreads x iters calls of the reference extractor (one instance per std::thread)."""
    frames = np.ascontiguousarray(frames, np.uint8)
    tot = C.c_longlong(0)
    dt = lib().ref_extract_throughput(nfeatures, 1.2, 8, 20, 7, _ptr(frames), frames.shape[0], frames.shape[1],
                                      frames.shape[2], nthreads, iters, C.byref(tot))
    return dt, tot.value


# ---- oracle/_ref/libref_bow.so: the reference's vendored DBoW2 (FORB.cpp, BowVector.cpp, FeatureVector.cpp,
#      ScoringObject.cpp, DUtils, TemplatedVocabulary.h instantiated for FORB) compiled unmodified against oracle/cvcompat/
BOW_LIB_PATH = os.path.join(_HERE, "_ref", "libref_bow.so")
_bow = None


def build_bow(force=False):
    if not os.path.exists(os.path.join(REFERENCE, "Thirdparty", "DBoW2", "DBoW2", "FORB.cpp")):
        return None
    cmd = ["make", "-C", _HERE, "REF=" + REFERENCE] + (["-B"] if force else []) + ["_ref/libref_bow.so"]
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    return BOW_LIB_PATH


def bow_available():
    return os.path.exists(BOW_LIB_PATH) or build_bow() is not None


def bow_lib():
    global _bow
    if _bow is None:
        if not os.path.exists(BOW_LIB_PATH) and build_bow() is None:
            raise FileNotFoundError("oracle/_ref/libref_bow.so is not built and %s is absent" % REFERENCE)
        L = C.CDLL(BOW_LIB_PATH)
        vp = C.c_void_p
        L.ref_voc_load_text.restype = vp
        L.ref_voc_load_text.argtypes = [C.c_char_p]
        L.ref_voc_free.argtypes = [vp]
        L.ref_voc_size.restype = C.c_uint
        L.ref_voc_size.argtypes = [vp]
        L.ref_voc_transform.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, C.POINTER(C.c_int), vp, vp, vp, C.POINTER(C.c_int)]
        _bow = L
    return _bow


def write_vocabulary_text(path, k, L, child_ptr, child_ids, desc, weight):
    """A flat vocabulary (the arrays of orb_vocab_view, node ids in creation order, parents before children) in the
    ORBvoc.txt format TemplatedVocabulary::loadFromTextFile reads (TemplatedVocabulary.h:1338-1419): `k L scoring
    weighting`, then one line per node `parent is_leaf d0 .. d31 weight`.  No trailing newline: the loader's
    `while(!f.eof())` would turn an empty last line into a node."""
    n = len(weight)
    parent = np.zeros(n, np.int64)
    for p in range(n):
        parent[child_ids[child_ptr[p]:child_ptr[p + 1]]] = p
    lines = ["%d %d 0 0" % (k, L)]   # L1_NORM, TF_IDF: the ORB vocabulary's settings
    for i in range(1, n):
        leaf = child_ptr[i + 1] == child_ptr[i]
        lines.append("%d %d %s %s" % (parent[i], 1 if leaf else 0, " ".join(str(int(b)) for b in desc[i]), repr(float(weight[i]))))
    with open(path, "w") as f:
        f.write("\n".join(lines))


class RefVocabulary:
    def __init__(self, path):
        self._h = bow_lib().ref_voc_load_text(path.encode())
        if not self._h:
            raise RuntimeError("loadFromTextFile failed: " + path)

    def __del__(self):
        if getattr(self, "_h", None):
            bow_lib().ref_voc_free(self._h)
            self._h = None

    def size(self):
        return int(bow_lib().ref_voc_size(self._h))

    def transform(self, desc, levelsup=4):
        """DBoW2 transform as Frame::ComputeBoW calls it.  Returns dict(bow_ids, bow_vals, fv_node_ids, fv_ptr, fv_idx)."""
        desc = np.ascontiguousarray(desc, np.uint8)
        n = len(desc)
        cap = max(n, 1)
        ids, vals = np.zeros(cap, np.uint32), np.zeros(cap)
        fn, fp, fi = np.zeros(cap, np.uint32), np.zeros(cap + 1, np.int32), np.zeros(cap, np.int32)
        nw, nn = C.c_int(0), C.c_int(0)
        rc = bow_lib().ref_voc_transform(self._h, _p(desc), n, int(levelsup), cap, _p(ids), _p(vals), C.byref(nw), _p(fn), _p(fp),
                                         _p(fi), C.byref(nn))
        assert rc == 0
        return dict(bow_ids=ids[:nw.value].astype(np.int32), bow_vals=vals[:nw.value].copy(), fv_node_ids=fn[:nn.value].astype(np.int32),
                    fv_ptr=fp[:nn.value + 1].copy(), fv_idx=fi[:fp[nn.value]].copy())


# ---- oracle/_ref/libref_front.so: the reference's front-end object code -- ORBmatcher.cc, Frame.cc, MapPoint.cc and
#      Pinhole.cpp compiled unmodified (oracle/Makefile); real Frame / MapPoint objects filled from the flat views
FRONT_LIB_PATH = os.path.join(_HERE, "_ref", "libref_front.so")
_front = None


def build_front(force=False):
    if not os.path.exists(os.path.join(REFERENCE, "src", "ORBmatcher.cc")):
        return None
    _o.build()
    cmd = ["make", "-C", _HERE, "REF=" + REFERENCE] + (["-B"] if force else []) + ["_ref/libref_front.so"]
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    return FRONT_LIB_PATH


def front_available():
    return os.path.exists(FRONT_LIB_PATH) or build_front() is not None


def front_lib():
    global _front
    if _front is None:
        if not os.path.exists(FRONT_LIB_PATH) and build_front() is None:
            raise FileNotFoundError("oracle/_ref/libref_front.so is not built and %s is absent" % REFERENCE)
        _o.lib()
        L = C.CDLL(FRONT_LIB_PATH)
        vp = C.c_void_p
        L.ref_front_project_local.argtypes = [vp, vp, C.c_float, C.c_float, C.c_int, C.c_float, vp]
        L.ref_front_project_last.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float, vp]
        L.ref_front_is_in_frustum.argtypes = [vp, C.c_float, vp, vp, vp, vp, vp, vp, vp]
        L.ref_front_stereo_match.argtypes = [C.c_int, vp, vp, C.c_int, vp, vp, C.c_int, vp, vp, vp, vp, vp, C.c_float, C.c_float,
                                             C.c_float, vp, vp]
        L.ref_front_triangulate.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp, vp]
        _front = L
    return _front


def front_project_local(F, mps, th, nn_ratio, far_points=False, th_far=50.0):
    """The reference's ORBmatcher(nn_ratio).SearchByProjection(Frame&, vector<MapPoint*>&, th, bFarPoints, thFarPoints)."""
    out = np.empty(F.n, np.int32)
    n = front_lib().ref_front_project_local(C.byref(F), C.byref(mps), th, nn_ratio, int(far_points), th_far, _p(out))
    return n, out


def front_project_last(cur, last, Tcw_qt7, th, forward=False, backward=False, check_ori=True, nn_ratio=0.9):
    """The reference's ORBmatcher(nn_ratio, checkOri).SearchByProjection(Cur, Last, th, bMono)."""
    T = np.ascontiguousarray(Tcw_qt7, np.float32)
    out = np.empty(cur.n, np.int32)
    n = front_lib().ref_front_project_last(C.byref(cur), C.byref(last), _p(T), int(forward), int(backward), th, int(check_ori),
                                           nn_ratio, _p(out))
    return n, out


def front_is_in_frustum(view, viewing_cos_limit=0.5, out=None):
    """The reference's Frame::isInFrustum over all points of an orb_frustum_view (outputs like oracle.is_in_frustum)."""
    if out is None:
        n = view.n
        out = dict(track_in_view=np.zeros(n, np.uint8), proj_x=np.zeros(n, np.float32), proj_y=np.zeros(n, np.float32),
                   proj_xr=np.zeros(n, np.float32), scale_level=np.zeros(n, np.int32),
                   view_cos=np.zeros(n, np.float32), depth=np.zeros(n, np.float32))
    k = front_lib().ref_front_is_in_frustum(C.byref(view), float(viewing_cos_limit),
                                            *[_p(out[f]) for f in ("track_in_view", "proj_x", "proj_y", "proj_xr", "scale_level",
                                                                   "view_cos", "depth")])
    return k, out


def front_stereo_match(kl, dl, kr, dr, pyr_l, pyr_r, bf, b, scale_factor=1.2):
    """The reference's Frame::ComputeStereoMatches (arguments like oracle.stereo_match).  Returns (n_kept, mvuRight, mvDepth)."""
    nl = len(pyr_l)
    Lv = [np.ascontiguousarray(a, np.uint8) for a in pyr_l]
    Rv = [np.ascontiguousarray(a, np.uint8) for a in pyr_r]
    pl = (C.c_void_p * nl)(*[a.ctypes.data for a in Lv])
    pr = (C.c_void_p * nl)(*[a.ctypes.data for a in Rv])
    lw = np.array([a.shape[1] for a in Lv], np.int32)
    lh = np.array([a.shape[0] for a in Lv], np.int32)
    ls = np.array([a.strides[0] for a in Lv], np.int32)
    kl = np.ascontiguousarray(kl); kr = np.ascontiguousarray(kr)
    dl = np.ascontiguousarray(dl, np.uint8); dr = np.ascontiguousarray(dr, np.uint8)
    ur = np.zeros(len(kl), np.float32)
    dp = np.zeros(len(kl), np.float32)
    n = front_lib().ref_front_stereo_match(len(kl), _p(kl), _p(dl), len(kr), _p(kr), _p(dr), nl, pl, pr, _p(lw), _p(lh), _p(ls),
                                           float(scale_factor), float(bf), float(b), _p(ur), _p(dp))
    return n, ur, dp


def front_triangulate(kf1, kf2, fv1, fv2, T1w_qt7, T2w_qt7, only_stereo=False, coarse=False, check_ori=True, cap=None):
    """The reference's ORBmatcher(0.6, check_ori).SearchForTriangulation(pKF1, pKF2, ...) on KeyFrames at the two poses.
    Returns (n, pairs, F12, ep): F12 / ep are what the reference derives from the poses (Pinhole.cpp:107-112, ORBmatcher.cc:917-920),
    the inputs of oracle.match_triangulate / the C ABI."""
    cap = cap or max(kf1.n, 1)
    T1 = np.ascontiguousarray(T1w_qt7, np.float32)
    T2 = np.ascontiguousarray(T2w_qt7, np.float32)
    out = np.empty((cap, 2), np.int32)
    F12 = np.zeros(9, np.float32)
    ep = np.zeros(2, np.float32)
    n = front_lib().ref_front_triangulate(C.byref(kf1), C.byref(kf2), C.byref(fv1), C.byref(fv2), _p(T1), _p(T2), int(only_stereo),
                                          int(coarse), int(check_ori), _p(out), cap, _p(F12), _p(ep))
    return n, out[:n], F12, ep


# ---- oracle/_ref/libref_g2o.so: the vendored g2o's types_six_dof_expmap.{h,cpp} as object code (stereo edges)
G2O_LIB_PATH = os.path.join(_HERE, "_ref", "libref_g2o.so")
_g2o = None


def build_g2o(force=False):
    if not os.path.exists(os.path.join(REFERENCE, "Thirdparty", "g2o", "g2o", "types", "types_six_dof_expmap.cpp")):
        return None
    cmd = ["make", "-C", _HERE, "REF=" + REFERENCE] + (["-B"] if force else []) + ["_ref/libref_g2o.so"]
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    return G2O_LIB_PATH


def g2o_available():
    return os.path.exists(G2O_LIB_PATH) or build_g2o() is not None


def g2o_lib():
    global _g2o
    if _g2o is None:
        if not os.path.exists(G2O_LIB_PATH) and build_g2o() is None:
            raise FileNotFoundError("oracle/_ref/libref_g2o.so is not built and %s is absent" % REFERENCE)
        L = C.CDLL(G2O_LIB_PATH)
        vp = C.c_void_p
        L.ref_g2o_edge_binary.argtypes = [C.c_int, vp, vp, vp, vp, C.c_double, vp, vp, vp, vp, vp]
        L.ref_g2o_edge_unary.argtypes = [C.c_int, vp, vp, vp, vp, C.c_double, vp, vp, vp, vp]
        L.ref_g2o_huber.argtypes = [C.c_double, C.c_double, vp]
        L.ref_g2o_oplus.argtypes = [vp, vp, vp]
        L.ref_g2o_build_system.argtypes = [vp, vp, vp, vp, vp, vp]
        L.ref_g2o_pose_system.argtypes = [vp, vp, vp]
        L.ref_g2o_se3_map.argtypes = [vp, vp, vp]
        _g2o = L
    return _g2o


def g2o_edge_binary(stereo, k5, pose7, X, obs, info=1.0):
    """g2o::EdgeStereoSE3ProjectXYZ (stereo) / g2o::EdgeSE3ProjectXYZ: err[d], Jxi[d,3] (point), Jxj[d,6] (pose),
    isDepthPositive, chi2 = e' (info I) e; d = 3 / 2.  k5 = fx fy cx cy bf."""
    d = 3 if stereo else 2
    k5, pose7, X, obs = (np.ascontiguousarray(a, np.float64) for a in (k5, pose7, X, obs))
    err, Jxi, Jxj, dp, chi = np.zeros(d), np.zeros((d, 3)), np.zeros((d, 6)), C.c_int(0), C.c_double(0)
    g2o_lib().ref_g2o_edge_binary(int(stereo), _p(k5), _p(pose7), _p(X), _p(obs), float(info), _p(err), _p(Jxi), _p(Jxj),
                                  C.byref(dp), C.byref(chi))
    return err, Jxi, Jxj, bool(dp.value), chi.value


def g2o_edge_unary(stereo, k5, pose7, Xw, obs, info=1.0):
    """g2o::EdgeStereoSE3ProjectXYZOnlyPose / g2o::EdgeSE3ProjectXYZOnlyPose: err[d], Jxi[d,6], isDepthPositive, chi2."""
    d = 3 if stereo else 2
    k5, pose7, Xw, obs = (np.ascontiguousarray(a, np.float64) for a in (k5, pose7, Xw, obs))
    err, J, dp, chi = np.zeros(d), np.zeros((d, 6)), C.c_int(0), C.c_double(0)
    g2o_lib().ref_g2o_edge_unary(int(stereo), _p(k5), _p(pose7), _p(Xw), _p(obs), float(info), _p(err), _p(J), C.byref(dp),
                                 C.byref(chi))
    return err, J, bool(dp.value), chi.value


# ---- oracle/_ref/libref_lm.so: g2o's optimization_algorithm_levenberg.cpp as object code over the oracle's Stepper
LM_LIB_PATH = os.path.join(_HERE, "_ref", "libref_lm.so")
_lm = None


def build_lm(force=False):
    if not os.path.exists(os.path.join(REFERENCE, "Thirdparty", "g2o", "g2o", "core", "optimization_algorithm_levenberg.cpp")):
        return None
    _o.build()
    cmd = ["make", "-C", _HERE, "REF=" + REFERENCE] + (["-B"] if force else []) + ["_ref/libref_lm.so"]
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    return LM_LIB_PATH


def lm_available():
    return os.path.exists(LM_LIB_PATH) or build_lm() is not None


def lm_lib():
    global _lm
    if _lm is None:
        if not os.path.exists(LM_LIB_PATH) and build_lm() is None:
            raise FileNotFoundError("oracle/_ref/libref_lm.so is not built and %s is absent" % REFERENCE)
        _lm = C.CDLL(LM_LIB_PATH)
    return _lm


def lm_driver():
    """Address of ref_lm_driver (an orc_lm_driver, oracle/orc_lm_ops.h): optimizer.optimize() run by the reference's
    OptimizationAlgorithmLevenberg object code.  Pass it as `driver=` to oracle.lba_solve / pose_optimize / lia_solve."""
    return C.cast(lm_lib().ref_lm_driver, C.c_void_p).value


def g2o_huber(delta, e):
    """g2o::RobustKernelHuber: setDelta(delta), robustify(e) -> rho[3] (rho, rho', rho'')."""
    rho = np.zeros(3)
    g2o_lib().ref_g2o_huber(float(delta), float(e), _p(rho))
    return rho


def g2o_oplus(pose7, update6):
    """g2o::VertexSE3Expmap::oplus(update6) on the estimate pose7, with the reference's own se3quat.h: the new pose7."""
    out = np.zeros(7)
    g2o_lib().ref_g2o_oplus(_p(np.ascontiguousarray(pose7, np.float64)), _p(np.ascontiguousarray(update6, np.float64)), _p(out))
    return out


def g2o_se3_map(pose7, X):
    out = np.zeros(3)
    g2o_lib().ref_g2o_se3_map(_p(np.ascontiguousarray(pose7, np.float64)), _p(np.ascontiguousarray(X, np.float64)), _p(out))
    return out


def g2o_build_system(g):
    """computeError + linearizeOplus + constructQuadraticForm of every edge of a Pinhole window, in edge order, as the
    vendored g2o's object code: dict like oracle.lba_system."""
    out = dict(Hpp=np.zeros((g.n_kf, 6, 6)), Hll=np.zeros((g.n_mp, 3, 3)), W=np.zeros((g.n_edges, 6, 3)),
               bp=np.zeros((g.n_kf, 6)), bl=np.zeros((g.n_mp, 3)))
    rc = g2o_lib().ref_g2o_build_system(C.byref(g), *[_p(out[k]) for k in ("Hpp", "Hll", "W", "bp", "bl")])
    if rc != 0:
        raise ValueError("second-camera edges are not g2o types")
    return out


def g2o_pose_system(view):
    """(H[6,6], b[6]): BaseUnaryEdge::constructQuadraticForm over the pose-only edges of a pose_opt_view (object code)."""
    H, b = np.zeros((6, 6)), np.zeros(6)
    g2o_lib().ref_g2o_pose_system(C.byref(view), _p(H), _p(b))
    return H, b


# ---- oracle/_ref/libref_vi.so: src/G2oTypes.cc as object code (the visual edges of LocalInertialBA)
VI_LIB_PATH = os.path.join(_HERE, "_ref", "libref_vi.so")
_vi = None


def build_vi(force=False):
    if not os.path.exists(os.path.join(REFERENCE, "src", "G2oTypes.cc")):
        return None
    cmd = ["make", "-C", _HERE, "REF=" + REFERENCE] + (["-B"] if force else []) + ["_ref/libref_vi.so"]
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    return VI_LIB_PATH


def vi_available():
    return os.path.exists(VI_LIB_PATH) or build_vi() is not None


def vi_edge(v, e):
    """The reference's EdgeMono / EdgeStereo (src/G2oTypes.cc) on edge e of a lia_graph_view at its input state:
    err[3], A[3,3] = d err / d point, B[3,6] = d err / d pose (third rows zero for EdgeMono), isDepthPositive."""
    global _vi
    if _vi is None:
        if not os.path.exists(VI_LIB_PATH) and build_vi() is None:
            raise FileNotFoundError("oracle/_ref/libref_vi.so is not built and %s is absent" % REFERENCE)
        _vi = C.CDLL(VI_LIB_PATH)
        _vi.ref_vi_edge.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    err, A, B, dp = np.zeros(3), np.zeros((3, 3)), np.zeros((3, 6)), C.c_int(0)
    rc = _vi.ref_vi_edge(C.byref(v), int(e), _p(err), _p(A), _p(B), C.byref(dp))
    assert rc == 0
    return err, A, B, bool(dp.value)

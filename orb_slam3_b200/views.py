"""ctypes mirrors of the flat view structs of include/orb_b200.h.

A view only borrows memory: every numpy array handed in is kept alive on the
Python object (`_keep`)."""
import ctypes as C

import numpy as np

from ._lib import KP_DTYPE

_f, _i = C.c_float, C.c_int32
_vp = C.c_void_p

FRAME_GRID_COLS, FRAME_GRID_ROWS = 64, 48  # include/Frame.h:44-45


class orb_frame_view(C.Structure):
    _fields_ = [("n", _i), ("keys", _vp), ("u_right", _vp), ("desc", _vp),
                ("min_x", _f), ("min_y", _f), ("max_x", _f), ("max_y", _f),
                ("grid_w_inv", _f), ("grid_h_inv", _f),
                ("n_levels", _i), ("scale_factors", _vp), ("level_sigma2", _vp),
                ("fx", _f), ("fy", _f), ("cx", _f), ("cy", _f), ("bf", _f), ("b", _f),
                ("kp_taken", _vp)]


class orb_mappoint_view(C.Structure):
    _fields_ = [("n", _i), ("track_in_view", _vp), ("is_bad", _vp), ("has_obs", _vp),
                ("proj_x", _vp), ("proj_y", _vp), ("proj_xr", _vp), ("scale_level", _vp),
                ("view_cos", _vp), ("depth", _vp), ("desc", _vp)]


class orb_lastframe_view(C.Structure):
    _fields_ = [("n", _i), ("has_mp", _vp), ("has_obs", _vp), ("world_pos", _vp), ("desc", _vp),
                ("octave", _vp), ("angle", _vp)]


class orb_featvec_view(C.Structure):
    _fields_ = [("n_nodes", _i), ("node_ids", _vp), ("ptr", _vp), ("idx", _vp)]


def _arr(a, dtype):
    if a is None:
        return None
    return np.ascontiguousarray(a, dtype=dtype)


def _p(a):
    return None if a is None else a.ctypes.data


def make_frame_view(keys, desc, width, height, scale_factors, u_right=None, kp_taken=None,
                    fx=700.0, fy=700.0, cx=None, cy=None, bf=0.0, b=0.0, level_sigma2=None):
    """Frame fields as the Frame constructor derives them for an undistorted pinhole
    image: mnMinX=0, mnMaxX=cols, grid inverse = 64/(maxX-minX) (Frame.cc:ComputeImageBounds,
    :181-182)."""
    keys = _arr(keys, KP_DTYPE)
    desc = _arr(desc, np.uint8)
    sf = _arr(scale_factors, np.float32)
    s2 = _arr(level_sigma2 if level_sigma2 is not None else sf * sf, np.float32)
    ur = _arr(u_right, np.float32)
    tk = _arr(kp_taken, np.uint8)
    v = orb_frame_view()
    v.n = len(keys)
    v.keys, v.u_right, v.desc, v.kp_taken = _p(keys), _p(ur), _p(desc), _p(tk)
    v.min_x, v.min_y, v.max_x, v.max_y = 0.0, 0.0, float(width), float(height)
    v.grid_w_inv = np.float32(FRAME_GRID_COLS) / np.float32(v.max_x - v.min_x)
    v.grid_h_inv = np.float32(FRAME_GRID_ROWS) / np.float32(v.max_y - v.min_y)
    v.n_levels = len(sf)
    v.scale_factors, v.level_sigma2 = _p(sf), _p(s2)
    v.fx, v.fy = fx, fy
    v.cx = width / 2.0 if cx is None else cx
    v.cy = height / 2.0 if cy is None else cy
    v.bf, v.b = bf, b
    v._keep = (keys, desc, sf, s2, ur, tk)
    return v


def make_mappoint_view(proj_x, proj_y, scale_level, desc, view_cos=None, proj_xr=None, depth=None,
                       track_in_view=None, is_bad=None, has_obs=None):
    n = len(proj_x)
    a = dict(
        track_in_view=_arr(np.ones(n) if track_in_view is None else track_in_view, np.uint8),
        is_bad=_arr(np.zeros(n) if is_bad is None else is_bad, np.uint8),
        has_obs=_arr(np.ones(n) if has_obs is None else has_obs, np.uint8),
        proj_x=_arr(proj_x, np.float32), proj_y=_arr(proj_y, np.float32),
        proj_xr=_arr(np.full(n, -1.0) if proj_xr is None else proj_xr, np.float32),
        scale_level=_arr(scale_level, np.int32),
        view_cos=_arr(np.ones(n) if view_cos is None else view_cos, np.float32),
        depth=_arr(np.ones(n) if depth is None else depth, np.float32),
        desc=_arr(desc, np.uint8))
    v = orb_mappoint_view()
    v.n = n
    for k, arr in a.items():
        setattr(v, k, _p(arr))
    v._keep = a
    return v


def make_lastframe_view(world_pos, desc, octave, angle, has_mp=None, has_obs=None):
    n = len(octave)
    a = dict(has_mp=_arr(np.ones(n) if has_mp is None else has_mp, np.uint8),
             has_obs=_arr(np.ones(n) if has_obs is None else has_obs, np.uint8),
             world_pos=_arr(world_pos, np.float32), desc=_arr(desc, np.uint8),
             octave=_arr(octave, np.int32), angle=_arr(angle, np.float32))
    v = orb_lastframe_view()
    v.n = n
    for k, arr in a.items():
        setattr(v, k, _p(arr))
    v._keep = a
    return v


def make_featvec_view(node_of_feature):
    """DBoW2::FeatureVector from a per-feature node id (-1 = feature not in the vector)."""
    node_of_feature = np.asarray(node_of_feature, dtype=np.int64)
    valid = np.nonzero(node_of_feature >= 0)[0]
    order = valid[np.argsort(node_of_feature[valid], kind="stable")]  # ascending index inside a node
    ids, counts = np.unique(node_of_feature[order], return_counts=True)
    a = dict(node_ids=_arr(ids, np.uint32), ptr=_arr(np.concatenate([[0], np.cumsum(counts)]), np.int32),
             idx=_arr(order, np.int32))
    v = orb_featvec_view()
    v.n_nodes = len(ids)
    v.node_ids, v.ptr, v.idx = _p(a["node_ids"]), _p(a["ptr"]), _p(a["idx"])
    v._keep = a
    return v


class lba_graph_view(C.Structure):
    _fields_ = [("n_kf", _i), ("kf_pose", _vp), ("kf_fixed", _vp), ("kf_cam", _vp),
                ("n_mp", _i), ("mp_pos", _vp),
                ("n_edges", _i), ("e_kf", _vp), ("e_mp", _vp), ("e_stereo", _vp), ("e_obs", _vp),
                ("e_inv_sigma2", _vp),
                ("kf_cam_model", _vp), ("kf_cam_dist", _vp), ("kf_cam2_model", _vp), ("kf_cam2", _vp), ("kf_trl", _vp)]


class lba_stats(C.Structure):
    _fields_ = [("iterations", _i), ("trials", _i), ("stopped", _i),
                ("chi2_initial", C.c_double), ("chi2_final", C.c_double), ("lambda_final", C.c_double),
                ("ms_total", C.c_double), ("ms_linearize", C.c_double), ("ms_schur", C.c_double),
                ("ms_solve", C.c_double), ("ms_update", C.c_double),
                ("n_free_kf", _i), ("n_pairs", _i), ("schur_flops", C.c_double),
                ("solver_kind", _i), ("envelope_rows_max", _i), ("ms_host_prep", C.c_double), ("ms_wall", C.c_double),
                ("allreduce_bytes_per_trial", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def make_lba_graph_view(kf_pose, kf_fixed, kf_cam, mp_pos, e_kf, e_mp, e_stereo, e_obs, e_inv_sigma2,
                        kf_cam_model=None, kf_cam_dist=None, kf_cam2_model=None, kf_cam2=None, kf_trl=None):
    a = dict(kf_pose=_arr(kf_pose, np.float64), kf_fixed=_arr(kf_fixed, np.uint8), kf_cam=_arr(kf_cam, np.float32),
             mp_pos=_arr(mp_pos, np.float64), e_kf=_arr(e_kf, np.int32), e_mp=_arr(e_mp, np.int32),
             e_stereo=_arr(e_stereo, np.uint8), e_obs=_arr(e_obs, np.float64),
             e_inv_sigma2=_arr(e_inv_sigma2, np.float32))
    # rig extension (include/orb_b200.h): absent = NULL = one Pinhole camera per keyframe
    for name, val, dt in (("kf_cam_model", kf_cam_model, np.uint8), ("kf_cam_dist", kf_cam_dist, np.float32),
                          ("kf_cam2_model", kf_cam2_model, np.uint8), ("kf_cam2", kf_cam2, np.float32),
                          ("kf_trl", kf_trl, np.float64)):
        if val is not None:
            a[name] = _arr(val, dt)
    v = lba_graph_view()
    v.n_kf, v.n_mp, v.n_edges = len(a["kf_fixed"]), len(a["mp_pos"]), len(a["e_kf"])
    for k, arr in a.items():
        setattr(v, k, _p(arr))
    v._keep = a
    return v


class pose_opt_view(C.Structure):
    _fields_ = [("n", _i), ("xw", _vp), ("obs", _vp), ("inv_sigma2", _vp),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("bf", C.c_float),
                ("pose", C.c_double * 7)]


def make_pose_opt_view(xw, obs, inv_sigma2, cam, pose):
    """xw [n,3] f32 world points, obs [n,3] f32 (u, v, uRight or -1), inv_sigma2 [n] f32,
    cam = (fx, fy, cx, cy, bf), pose = quaternion xyzw + translation."""
    a = dict(xw=_arr(xw, np.float32), obs=_arr(obs, np.float32), inv_sigma2=_arr(inv_sigma2, np.float32))
    v = pose_opt_view()
    v.n = len(a["inv_sigma2"])
    for k, arr in a.items():
        setattr(v, k, _p(arr))
    v.fx, v.fy, v.cx, v.cy, v.bf = (float(c) for c in cam)
    for i in range(7):
        v.pose[i] = float(pose[i])
    v._keep = a
    return v


class orb_frustum_view(C.Structure):
    _fields_ = [("n", _i), ("world_pos", _vp), ("normal", _vp), ("min_dist", _vp), ("max_dist", _vp),
                ("Rcw", C.c_float * 9), ("tcw", C.c_float * 3), ("Ow", C.c_float * 3),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("bf", C.c_float),
                ("min_x", C.c_float), ("max_x", C.c_float), ("min_y", C.c_float), ("max_y", C.c_float),
                ("log_scale_factor", C.c_float), ("n_levels", _i)]


def make_frustum_view(world_pos, normal, min_dist, max_dist, Rcw, tcw, cam, bounds, scale_factor=1.2, n_levels=8):
    """Frame side: Rcw [3,3], tcw [3] (Ow = -Rcw^T tcw in float like Frame::UpdatePoseMatrices),
    cam = (fx, fy, cx, cy, bf), bounds = (min_x, max_x, min_y, max_y)."""
    a = dict(world_pos=_arr(world_pos, np.float32), normal=_arr(normal, np.float32),
             min_dist=_arr(min_dist, np.float32), max_dist=_arr(max_dist, np.float32))
    v = orb_frustum_view()
    v.n = len(a["min_dist"])
    for k, arr in a.items():
        setattr(v, k, _p(arr))
    R = np.asarray(Rcw, np.float32).reshape(3, 3)
    t = np.asarray(tcw, np.float32).reshape(3)
    Ow = (-(R.T @ t)).astype(np.float32)
    for i in range(9):
        v.Rcw[i] = float(R.reshape(9)[i])
    for i in range(3):
        v.tcw[i] = float(t[i])
        v.Ow[i] = float(Ow[i])
    v.fx, v.fy, v.cx, v.cy, v.bf = (float(c) for c in cam)
    v.min_x, v.max_x, v.min_y, v.max_y = (float(b) for b in bounds)
    v.log_scale_factor = float(np.log(np.float32(scale_factor)))  # std::log(float) (Frame.cc:112)
    v.n_levels = int(n_levels)
    v._keep = a
    return v


class orb_vocab_view(C.Structure):
    _fields_ = [("n_nodes", _i), ("L", _i), ("child_ptr", _vp), ("child_ids", _vp), ("desc", _vp), ("weight", _vp),
                ("word_id", _vp)]


def make_vocab_view(L, child_ptr, child_ids, desc, weight, word_id):
    a = dict(child_ptr=_arr(child_ptr, np.int32), child_ids=_arr(child_ids, np.int32), desc=_arr(desc, np.uint8),
             weight=_arr(weight, np.float64), word_id=_arr(word_id, np.int32))
    v = orb_vocab_view()
    v.n_nodes, v.L = len(a["weight"]), int(L)
    for k, arr in a.items():
        setattr(v, k, _p(arr))
    v._keep = a
    return v


# ---- Optimizer::LocalInertialBA graph (include/orb_b200.h: lia_graph_view)
class lia_graph_view(C.Structure):
    _fields_ = [("n_kf", _i), ("kf_Rwb", _vp), ("kf_twb", _vp), ("kf_Rcw", _vp), ("kf_tcw", _vp), ("kf_fixed", _vp),
                ("kf_has_imu", _vp), ("kf_vel", _vp), ("kf_bg", _vp), ("kf_ba", _vp),
                ("Rcb", C.c_double * 9), ("tcb", C.c_double * 3), ("tbc", C.c_double * 3),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("bf", C.c_float),
                ("n_mp", _i), ("mp_pos", _vp),
                ("n_edges", _i), ("e_kf", _vp), ("e_mp", _vp), ("e_stereo", _vp), ("e_obs", _vp), ("e_inv_sigma2", _vp),
                ("n_inertial", _i), ("i_kf1", _vp), ("i_kf2", _vp), ("i_dR", _vp), ("i_dV", _vp), ("i_dP", _vp),
                ("i_JRg", _vp), ("i_JVg", _vp), ("i_JVa", _vp), ("i_JPg", _vp), ("i_JPa", _vp), ("i_bias", _vp),
                ("i_dT", _vp), ("i_C", _vp), ("i_last", _vp),
                ("lambda_init", C.c_double), ("iterations", _i)]


_LIA_TYPES = dict(kf_Rwb=np.float64, kf_twb=np.float64, kf_Rcw=np.float64, kf_tcw=np.float64, kf_fixed=np.uint8,
                  kf_has_imu=np.uint8, kf_vel=np.float64, kf_bg=np.float64, kf_ba=np.float64, mp_pos=np.float64,
                  e_kf=np.int32, e_mp=np.int32, e_stereo=np.uint8, e_obs=np.float64, e_inv_sigma2=np.float32,
                  i_kf1=np.int32, i_kf2=np.int32, i_dR=np.float32, i_dV=np.float32, i_dP=np.float32, i_JRg=np.float32,
                  i_JVg=np.float32, i_JVa=np.float32, i_JPg=np.float32, i_JPa=np.float32, i_bias=np.float32,
                  i_dT=np.float32, i_C=np.float32, i_last=np.uint8)


def make_lia_view(d):
    """d: dict with the arrays of lia_graph_view plus Rcb[3,3], tcb[3], tbc[3], cam=(fx,fy,cx,cy,bf),
    lambda_init, iterations."""
    v = lia_graph_view()
    keep = {}
    for k, t in _LIA_TYPES.items():
        keep[k] = np.ascontiguousarray(d[k], t)
        setattr(v, k, keep[k].ctypes.data)
    v.n_kf, v.n_mp, v.n_edges, v.n_inertial = len(keep["kf_fixed"]), len(keep["mp_pos"]), len(keep["e_kf"]), len(keep["i_kf1"])
    for i, x in enumerate(np.asarray(d["Rcb"], np.float64).reshape(9)):
        v.Rcb[i] = x
    for i in range(3):
        v.tcb[i] = float(d["tcb"][i])
        v.tbc[i] = float(d["tbc"][i])
    v.fx, v.fy, v.cx, v.cy, v.bf = (float(c) for c in d["cam"])
    v.lambda_init, v.iterations = float(d.get("lambda_init", 1.0)), int(d.get("iterations", 10))
    v._keep = keep
    return v

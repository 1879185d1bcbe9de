"""Host-side mirror of Optimizer::LocalBundleAdjustment's optimisation core
(src/Optimizer.cc:1116-1498, the part between graph set-up and write-back) over
the C ABI: lba_solve runs g2o's `optimizer.optimize(10)` equivalent on the GPU.
No CPU fallback."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, ptr
from .views import lba_stats


class LocalBundleAdjustment:
    """One solver handle (CUDA stream + scratch) reusable across calls."""

    CHI2_MONO = 5.991     # Optimizer.cc:1423
    CHI2_STEREO = 7.815   # Optimizer.cc:1455

    def __init__(self, device=0):
        self._lib = _lib.lib()
        h = C.c_void_p()
        check(self._lib.lba_create(int(device), C.byref(h)))
        self._h = h
        self.rank, self.world = 0, 1

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.lba_destroy(h)
            self._h = None

    def init_comm(self, rank, world, unique_id):
        """Landmark-sharded multi-GPU mode; unique_id = bytes from nccl_unique_id() of rank 0."""
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        check(self._lib.lba_comm_init(self._h, int(rank), int(world), buf))
        self.rank, self.world = rank, world

    @staticmethod
    def nccl_unique_id():
        buf = (C.c_char * 128)()
        check(_lib.lib().lba_nccl_unique_id(buf))
        return bytes(buf)

    def __call__(self, graph, pbStopFlag=None, max_iters=10, lambda_init=0.0):
        """graph: lba_graph_view.  Returns dict(iterations, kf_pose, mp_pos, chi2, depth_pos, stats);
        the caller applies the chi2 / depth tests (:1413-1460) and writes poses back."""
        kf = np.empty((graph.n_kf, 7))
        mp = np.empty((graph.n_mp, 3))
        chi2 = np.empty(graph.n_edges)
        dp = np.empty(graph.n_edges, np.uint8)
        st = lba_stats()
        sp = ptr(pbStopFlag) if pbStopFlag is not None else None
        it = check(self._lib.lba_solve(self._h, C.byref(graph), sp, int(max_iters), float(lambda_init), ptr(kf),
                                       ptr(mp), ptr(chi2), ptr(dp), C.byref(st)))
        return dict(iterations=it, kf_pose=kf, mp_pos=mp, chi2=chi2, depth_pos=dp, stats=st.as_dict())

    def kernel_launches(self):
        return int(self._lib.lba_kernel_launches(self._h))

    @classmethod
    def outliers(cls, graph_dict, result):
        """Edges the reference erases after optimize(): chi2 > 5.991 (mono and second-camera edges, Optimizer.cc:1424,
        :1438) / 7.815 (stereo, :1453) or depth <= 0."""
        st = np.asarray(graph_dict["e_stereo"]) == 1
        th = np.where(st, cls.CHI2_STEREO, cls.CHI2_MONO)
        return (result["chi2"] > th) | (result["depth_pos"] == 0)


class PoseOptimization:
    """Optimizer::PoseOptimization (src/Optimizer.cc:814-1115) over the C ABI: one frame or a batch of
    independent frames per call, one CTA per frame in a single kernel launch.  No CPU fallback."""

    CHI2_MONO = 5.991
    CHI2_STEREO = 7.815

    def __init__(self, device=0):
        self._lib = _lib.lib()
        h = C.c_void_p()
        check(self._lib.poseopt_create(int(device), C.byref(h)))
        self._h = h

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.poseopt_destroy(h)
            self._h = None

    def __call__(self, view):
        """view: pose_opt_view.  Returns (nInitialCorrespondences - nBad, pose[7], mvbOutlier[n])."""
        pose = np.empty(7)
        out = np.empty(max(view.n, 1), np.uint8)
        n = check(self._lib.pose_optimize(self._h, C.byref(view), ptr(pose), ptr(out)))
        return n, pose, out[:view.n].astype(bool)

    def batch(self, views):
        """Returns (inliers[B], poses[B,7], [mvbOutlier_k], stats[B,3] = rounds, LM iterations, LM trials)."""
        from .views import pose_opt_view
        B = len(views)
        arr = (pose_opt_view * B)(*views)
        pose = np.empty((B, 7))
        outs = [np.empty(max(v.n, 1), np.uint8) for v in views]
        optr = (C.c_void_p * B)(*[o.ctypes.data for o in outs])
        inl = np.empty(B, np.int32)
        stats = np.empty((B, 3), np.int32)
        check(self._lib.pose_optimize_batch(self._h, B, arr, ptr(pose), optr, ptr(inl), ptr(stats)))
        return inl, pose, [o[:v.n].astype(bool) for o, v in zip(outs, views)], stats

    def last_ms(self):
        return float(self._lib.poseopt_last_ms(self._h))

    def kernel_launches(self):
        return int(self._lib.poseopt_kernel_launches(self._h))


def _lia_outputs(view):
    kf = np.zeros((view.n_kf, 21))
    mp = np.zeros((max(view.n_mp, 1), 3))
    chi2 = np.zeros(max(view.n_edges, 1))
    dp = np.zeros(max(view.n_edges, 1), np.uint8)
    st = np.zeros(8)
    return kf, mp, chi2, dp, st


def _lia_pack(view, it, kf, mp, chi2, dp, st):
    return dict(iterations=it, Rcw=kf[:, :9].reshape(-1, 3, 3), tcw=kf[:, 9:12], vel=kf[:, 12:15], bg=kf[:, 15:18],
                ba=kf[:, 18:21], mp_pos=mp[:view.n_mp], chi2=chi2[:view.n_edges], depth_pos=dp[:view.n_edges],
                stats=dict(iterations=int(st[0]), trials=int(st[1]), err=st[2], err_end=st[3], lambda_final=st[4],
                           dim=int(st[5])))


class LocalInertialBA:
    """Optimizer::LocalInertialBA's optimizer.optimize(opt_it) (src/Optimizer.cc:2383-2958) over the C ABI: the whole
    LM loop in one kernel launch.  No CPU fallback (`lia_debug_host` runs the kernel's source on the host for the
    CPU tests)."""

    def __init__(self, device=0):
        self._lib = _lib.lib()
        h = C.c_void_p()
        check(self._lib.lia_create(int(device), C.byref(h)))
        self._h = h

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.lia_destroy(h)
            self._h = None

    def __call__(self, view):
        kf, mp, chi2, dp, st = _lia_outputs(view)
        it = check(self._lib.lia_solve(self._h, C.byref(view), ptr(kf), ptr(mp), ptr(chi2), ptr(dp), ptr(st)))
        return _lia_pack(view, it, kf, mp, chi2, dp, st)

    def last_ms(self):
        return float(self._lib.lia_last_ms(self._h))

    def kernel_launches(self):
        return int(self._lib.lia_kernel_launches(self._h))


def lia_debug_host(view):
    """csrc/lia_core.h on the host (CPU tests): same source as the kernel."""
    kf, mp, chi2, dp, st = _lia_outputs(view)
    it = check(_lib.lib().lia_debug_host(C.byref(view), ptr(kf), ptr(mp), ptr(chi2), ptr(dp), ptr(st)))
    return _lia_pack(view, it, kf, mp, chi2, dp, st)

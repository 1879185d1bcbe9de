"""Host-side mirror of ORB_SLAM3::ORBextractor over the C ABI.

Same constructor arguments, getters and call semantics as the reference class
(include/ORBextractor.h:49-83, src/ORBextractor.cc:1086-1168); all computation
happens in liborbb200.so on the GPU.  No CPU fallback.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import KP_DTYPE, check, ptr


class ORBextractor:
    HARRIS_SCORE = 0
    FAST_SCORE = 1

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, device=0, max_batch=1):
        self._lib = _lib.lib()
        h = C.c_void_p()
        check(self._lib.orb_create(int(nfeatures), float(scaleFactor), int(nlevels), int(iniThFAST),
                                   int(minThFAST), int(device), C.byref(h)))
        self._h = h
        self.nfeatures = int(nfeatures)
        self.nlevels = int(nlevels)
        self.max_batch = max_batch
        # a little above any possible count: quota + 3 overshoot per level
        self.cap = self.nfeatures + 8 * self.nlevels + 64

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.orb_destroy(h)
            self._h = None

    # ---- getters (ORBextractor.h:61-81)
    def GetLevels(self):
        return self._lib.orb_get_levels(self._h)

    def GetScaleFactor(self):
        return self._lib.orb_get_scale_factor(self._h)

    def _tab(self, fn, dtype=np.float32):
        out = np.zeros(self.nlevels, dtype=dtype)
        check(fn(self._h, ptr(out)))
        return out

    def GetScaleFactors(self):
        return self._tab(self._lib.orb_get_scale_factors)

    def GetInverseScaleFactors(self):
        return self._tab(self._lib.orb_get_inverse_scale_factors)

    def GetScaleSigmaSquares(self):
        return self._tab(self._lib.orb_get_scale_sigma_squares)

    def GetInverseScaleSigmaSquares(self):
        return self._tab(self._lib.orb_get_inverse_scale_sigma_squares)

    def features_per_level(self):
        return self._tab(self._lib.orb_get_features_per_level, np.int32)

    # ---- operator()
    def __call__(self, image, mask=None, vLappingArea=(0, 0)):
        """Returns (monoIndex, keypoints[KP_DTYPE], descriptors[n,32] uint8).

        An empty image returns (-1, empty, empty) like the reference (:1090)."""
        if image is None or image.size == 0:
            return -1, np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        assert image.dtype == np.uint8 and image.ndim == 2, "CV_8UC1 expected (:1094)"
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        kps = np.empty(self.cap, dtype=KP_DTYPE)
        desc = np.empty((self.cap, 32), dtype=np.uint8)
        n = C.c_int(0)
        mono = check(self._lib.orb_extract(self._h, ptr(image), image.shape[0], image.shape[1],
                                           image.strides[0], int(vLappingArea[0]), int(vLappingArea[1]),
                                           ptr(kps), ptr(desc), self.cap, C.byref(n)))
        return mono, kps[:n.value], desc[:n.value]

    def extract_batch(self, images, lapping=None):
        """`images`: sequence of equally sized uint8 images (or a [B,H,W] array).
        Returns a list of (monoIndex, keypoints, descriptors)."""
        B = len(images)
        imgs = [np.ascontiguousarray(im, dtype=np.uint8) for im in images]
        rows, cols = imgs[0].shape
        assert all(im.shape == (rows, cols) for im in imgs)
        arr = (C.c_void_p * B)(*[im.ctypes.data for im in imgs])
        kps = np.empty((B, self.cap), dtype=KP_DTYPE)
        desc = np.empty((B, self.cap, 32), dtype=np.uint8)
        n = np.zeros(B, np.int32)
        mono = np.zeros(B, np.int32)
        lap = None
        if lapping is not None:
            lap = np.ascontiguousarray(lapping, dtype=np.int32).reshape(B, 2)
        check(self._lib.orb_extract_batch(self._h, B, arr, rows, cols, imgs[0].strides[0],
                                          ptr(lap) if lap is not None else None, ptr(kps), ptr(desc),
                                          self.cap, ptr(n), ptr(mono)))
        return [(int(mono[b]), kps[b, :n[b]], desc[b, :n[b]]) for b in range(B)]

    def extract_batch_device(self, d_ptr, batch, rows, cols, step, frame_stride, lapping=None, stream=None):
        """Frames already resident in HBM (raw device pointer, e.g. tensor.data_ptr())."""
        lap = None
        if lapping is not None:
            lap = np.ascontiguousarray(lapping, dtype=np.int32).reshape(batch, 2)
        check(self._lib.orb_extract_batch_device(self._h, batch, C.c_void_p(d_ptr), frame_stride, rows, cols,
                                                 step, ptr(lap) if lap is not None else None,
                                                 C.c_void_p(stream) if stream else None))

    def device_results(self):
        k, d, n, m = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        cap = C.c_int()
        check(self._lib.orb_device_results(self._h, C.byref(k), C.byref(d), C.byref(n), C.byref(m),
                                           C.byref(cap)))
        return k.value, d.value, n.value, m.value, cap.value

    def download_results(self, frame):
        kps = np.empty(self.cap, dtype=KP_DTYPE)
        desc = np.empty((self.cap, 32), dtype=np.uint8)
        n = C.c_int(0)
        mono = check(self._lib.orb_download_results(self._h, frame, ptr(kps), ptr(desc), self.cap, C.byref(n)))
        return mono, kps[:n.value], desc[:n.value]

    def synchronize(self):
        check(self._lib.orb_synchronize(self._h))

    # ---- mvImagePyramid (ORBextractor.h:83)
    def image_pyramid(self, level, frame=0):
        p = C.c_void_p()
        r, c = C.c_int(), C.c_int()
        st = C.c_size_t()
        check(self._lib.orb_pyramid(self._h, frame, level, C.byref(p), C.byref(r), C.byref(c), C.byref(st)))
        buf = (C.c_uint8 * (st.value * r.value)).from_address(p.value)
        a = np.frombuffer(buf, dtype=np.uint8).reshape(r.value, st.value)[:, :c.value]
        return a.copy()

    # ---- instrumentation
    def set_profiling(self, on):
        check(self._lib.orb_set_profiling(self._h, int(on)))

    def stage_times(self, reset=True):
        ms = np.zeros(8, np.float64)
        ln = np.zeros(8, np.int64)
        check(self._lib.orb_stage_times(self._h, ptr(ms), ptr(ln), int(reset)))
        names = [self._lib.orb_stage_name(i).decode() for i in range(8)]
        return {nm: (float(ms[i]), int(ln[i])) for i, nm in enumerate(names)}

    def kernel_launches(self):
        return int(self._lib.orb_kernel_launches(self._h))

    def debug_candidates(self, level, frame=0, cap=400000):
        out = np.zeros((cap, 3), np.int32)
        n = check(self._lib.orb_debug_candidates(self._h, frame, level, ptr(out), cap))
        return out[:min(n, cap)].copy()

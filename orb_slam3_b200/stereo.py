"""Host-side mirror of Frame::ComputeStereoMatches (src/Frame.cc:811-981) over the C ABI.

The reference method reads the two extractors' keypoints, descriptors and
mvImagePyramid; here the same data is taken from what the two ORBextractor
handles left on the device after their last extract.  No CPU fallback."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, ptr


class StereoMatcher:
    def __init__(self, device=0):
        self._lib = _lib.lib()
        h = C.c_void_p()
        check(self._lib.stereo_create(int(device), C.byref(h)))
        self._h = h

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.stereo_destroy(h)
            self._h = None

    def ComputeStereoMatches(self, extractorLeft, extractorRight, n_left, mbf, mb):
        """One pair: returns (n_matched, mvuRight[n_left], mvDepth[n_left])."""
        cap = extractorLeft.cap
        ur = np.empty(cap, np.float32)
        dp = np.empty(cap, np.float32)
        n = check(self._lib.stereo_match(self._h, extractorLeft._h, extractorRight._h, float(mbf), float(mb),
                                         ptr(ur), ptr(dp), cap))
        return n, ur[:n_left].copy(), dp[:n_left].copy()

    def compute_batch(self, extractorLeft, extractorRight, batch, mbf, mb, on_device=False, cuda_stream=None):
        """`batch` pairs (frame i of both handles' last batch).  Host mode returns
        (kept[batch], mvuRight[batch, cap], mvDepth[batch, cap]); device mode only enqueues."""
        s = C.c_void_p(cuda_stream) if cuda_stream else None
        if on_device:
            check(self._lib.stereo_match_batch(self._h, extractorLeft._h, extractorRight._h, int(batch), float(mbf),
                                               float(mb), None, None, 0, None, 1, s))
            return None
        cap = extractorLeft.cap
        ur = np.empty((batch, cap), np.float32)
        dp = np.empty((batch, cap), np.float32)
        kept = np.empty(batch, np.int32)
        check(self._lib.stereo_match_batch(self._h, extractorLeft._h, extractorRight._h, int(batch), float(mbf),
                                           float(mb), ptr(ur), ptr(dp), cap, ptr(kept), 0, s))
        return kept, ur, dp

    def last_ms(self):
        return float(self._lib.stereo_last_ms(self._h))

    def kernel_launches(self):
        return int(self._lib.stereo_kernel_launches(self._h))

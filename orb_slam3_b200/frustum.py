"""Host-side mirror of Frame::isInFrustum over all local map points of a frame (src/Frame.cc:512-570,
called from Tracking::SearchLocalPoints) over the C ABI.  No CPU fallback (`debug_host` runs the
kernel's per-point source on the host for the CPU tests only)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, ptr


def _outputs(n, prev=None):
    if prev is not None:
        return prev
    return dict(track_in_view=np.zeros(n, np.uint8), proj_x=np.zeros(n, np.float32), proj_y=np.zeros(n, np.float32),
                proj_xr=np.zeros(n, np.float32), scale_level=np.zeros(n, np.int32), view_cos=np.zeros(n, np.float32),
                depth=np.zeros(n, np.float32))


def _args(o):
    return [ptr(o[k]) for k in ("track_in_view", "proj_x", "proj_y", "proj_xr", "scale_level", "view_cos", "depth")]


class FrustumCuller:
    def __init__(self, device=0):
        self._lib = _lib.lib()
        h = C.c_void_p()
        check(self._lib.frustum_create(int(device), C.byref(h)))
        self._h = h

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.frustum_destroy(h)
            self._h = None

    def isInFrustum(self, view, viewingCosLimit=0.5, out=None):
        """All n map points at once.  `out` (dict of arrays from an earlier call) keeps the members the
        reference leaves stale for points that are not in view.  Returns (n_in_view, out)."""
        o = _outputs(view.n, out)
        n = check(self._lib.frame_is_in_frustum(self._h, C.byref(view), float(viewingCosLimit), *_args(o)))
        return n, o

    def enqueue(self, view, viewingCosLimit=0.5, cuda_stream=None):
        check(self._lib.frame_is_in_frustum_device(self._h, C.byref(view), float(viewingCosLimit),
                                                   C.c_void_p(cuda_stream) if cuda_stream else None))

    def device_results(self):
        p = [C.c_void_p() for _ in range(8)]
        check(self._lib.frustum_device_results(self._h, *[C.byref(x) for x in p]))
        keys = ("track_in_view", "proj_x", "proj_y", "proj_xr", "scale_level", "view_cos", "depth", "count")
        return dict(zip(keys, [x.value for x in p]))

    def last_ms(self):
        return float(self._lib.frustum_last_ms(self._h))

    def kernel_launches(self):
        return int(self._lib.frustum_kernel_launches(self._h))


def debug_host(view, viewingCosLimit=0.5, out=None):
    """frustum_core.h on the host (CPU tests): same source as the kernel body."""
    o = _outputs(view.n, out)
    n = check(_lib.lib().frustum_debug_host(C.byref(view), float(viewingCosLimit), *_args(o)))
    return n, o

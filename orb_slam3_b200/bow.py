"""Host-side mirror of ORBVocabulary::transform as Frame::ComputeBoW uses it (src/Frame.cc:738-745,
Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1195) over the C ABI.  The vocabulary is uploaded
once per handle.  No CPU fallback (`debug_host` runs the kernels' source on the host for the CPU tests)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, ptr


def _bufs(n):
    cap = max(n, 1)
    return (cap, np.zeros(cap, np.int32), np.zeros(cap), np.zeros(cap, np.int32), np.zeros(cap + 1, np.int32),
            np.zeros(cap, np.int32), C.c_int(0), C.c_int(0))


def _pack(used, ids, vals, fn, fp, fi, nw, nn):
    return dict(used=used, bow_ids=ids[:nw.value].copy(), bow_vals=vals[:nw.value].copy(),
                fv_node_ids=fn[:nn.value].copy(), fv_ptr=fp[:nn.value + 1].copy(), fv_idx=fi[:used].copy())


class ORBVocabulary:
    def __init__(self, vocab_view, device=0):
        self._lib = _lib.lib()
        self._view = vocab_view  # keeps the host arrays alive during the upload
        h = C.c_void_p()
        check(self._lib.vocab_create(int(device), C.byref(vocab_view), C.byref(h)))
        self._h = h

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.vocab_destroy(h)
            self._h = None

    def transform(self, descriptors, levelsup=4):
        """(BowVector, FeatureVector) flattened: dict(used, bow_ids, bow_vals, fv_node_ids, fv_ptr, fv_idx)."""
        d = np.ascontiguousarray(descriptors, np.uint8)
        cap, ids, vals, fn, fp, fi, nw, nn = _bufs(len(d))
        used = check(self._lib.bow_transform(self._h, ptr(d), len(d), int(levelsup), ptr(ids), ptr(vals), C.byref(nw),
                                             ptr(fn), ptr(fp), ptr(fi), C.byref(nn), cap))
        return _pack(used, ids, vals, fn, fp, fi, nw, nn)

    def transform_extracted(self, extractor, frame=0, levelsup=4):
        """Descriptors of `frame` of the extractor's last batch, read where they are (device)."""
        cap, ids, vals, fn, fp, fi, nw, nn = _bufs(extractor.cap)
        used = check(self._lib.bow_transform_extracted(self._h, extractor._h, int(frame), int(levelsup), ptr(ids),
                                                       ptr(vals), C.byref(nw), ptr(fn), ptr(fp), ptr(fi),
                                                       C.byref(nn), cap))
        return _pack(used, ids, vals, fn, fp, fi, nw, nn)

    def last_ms(self):
        return float(self._lib.bow_last_ms(self._h))

    def kernel_launches(self):
        return int(self._lib.bow_kernel_launches(self._h))


def debug_host(vocab_view, descriptors, levelsup=4):
    """bow_core.h on the host (CPU tests): same source as the kernels."""
    d = np.ascontiguousarray(descriptors, np.uint8)
    cap, ids, vals, fn, fp, fi, nw, nn = _bufs(len(d))
    used = check(_lib.lib().bow_debug_host(C.byref(vocab_view), ptr(d), len(d), int(levelsup), ptr(ids), ptr(vals),
                                           C.byref(nw), ptr(fn), ptr(fp), ptr(fi), C.byref(nn), cap))
    return _pack(used, ids, vals, fn, fp, fi, nw, nn)

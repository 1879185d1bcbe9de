"""CPU-only checks of the product's host logic: the C ABI loads and exports
every declared symbol, the libstdc++ introsort emulation and the array octree
formulation agree with the oracle (std::sort / list-based DistributeOctTree)."""
import math
import os
import re

import numpy as np
import pytest

from orb_slam3_b200 import _lib as L
from orb_slam3_b200.synth import synth_frame

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "orb_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b((?:orb|match|ham|lba)_[a-z0-9_]+)\s*\(", hdr))
    lib = L.lib()
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), name
    assert declared <= set(L.SIGNATURES), declared - set(L.SIGNATURES)


def test_tables_match_reference_constants():
    import ctypes as C
    lib = L.lib()
    h = C.c_void_p()
    assert lib.orb_create(2000, 1.2, 8, 20, 7, 0, C.byref(h)) == 0
    q = np.zeros(8, np.int32)
    lib.orb_get_features_per_level(h, L.ptr(q))
    assert list(q) == [434, 362, 302, 251, 209, 175, 145, 122]
    s = np.zeros(8, np.float32)
    lib.orb_get_scale_factors(h, L.ptr(s))
    ref = [np.float32(1.0)]
    for _ in range(7):
        ref.append(np.float32(np.float64(ref[-1]) * np.float64(np.float32(1.2))))
    assert np.array_equal(s, np.array(ref, np.float32))
    lib.orb_destroy(h)
    h = C.c_void_p()
    assert lib.orb_create(1000, 1.2, 8, 20, 7, 0, C.byref(h)) == 0
    lib.orb_get_features_per_level(h, L.ptr(q))
    assert list(q) == [217, 181, 151, 126, 105, 87, 73, 60]
    lib.orb_destroy(h)


def test_no_device_fails_loudly():
    import ctypes as C
    lib = L.lib()
    if lib.orb_device_count() > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert lib.orb_create(1000, 1.2, 8, 20, 7, 0, C.byref(h)) == 0
    img = synth_frame(480, 640, 1)
    kps = np.zeros(2000, L.KP_DTYPE)
    desc = np.zeros((2000, 32), np.uint8)
    n = C.c_int()
    rc = lib.orb_extract(h, L.ptr(img), 480, 640, 640, 0, 0, L.ptr(kps), L.ptr(desc), 2000, C.byref(n))
    assert rc == -5  # ORB_E_NODEVICE: no CPU path
    lib.orb_destroy(h)


def test_introsort_emulation_equals_std_sort(oracle):
    lib = L.lib()
    rng = np.random.default_rng(0)
    for t in range(600):
        n = int(rng.integers(0, 600)) if t % 10 else int(rng.integers(600, 5000))
        cnt = rng.integers(2, 2 + int(rng.integers(1, 8)), n).astype(np.int32)
        ulx = (rng.integers(0, int(rng.integers(1, 40)), n) * 13).astype(np.int32)
        b = np.zeros(n, np.int32)
        lib.orb_debug_introsort(L.ptr(cnt), L.ptr(ulx), n, L.ptr(b))
        assert np.array_equal(oracle.sort_nodes(cnt, ulx), b)
        lib.orb_debug_introsort_levels(L.ptr(cnt), L.ptr(ulx), n, L.ptr(b))      # the form the octree CTA runs
        assert np.array_equal(oracle.sort_nodes(cnt, ulx), b)
    for n in (17, 100, 1000, 5000):  # sorted / reversed / organ pipe / constant
        for cnt in (np.arange(n), np.arange(n)[::-1],
                    np.concatenate([np.arange(n // 2), np.arange(n - n // 2)[::-1]]), np.zeros(n)):
            cnt = np.ascontiguousarray(cnt, dtype=np.int32)
            ulx = np.zeros(n, np.int32)
            b = np.zeros(n, np.int32)
            lib.orb_debug_introsort(L.ptr(cnt), L.ptr(ulx), n, L.ptr(b))
            assert np.array_equal(oracle.sort_nodes(cnt, ulx), b)
            lib.orb_debug_introsort_levels(L.ptr(cnt), L.ptr(ulx), n, L.ptr(b))
            assert np.array_equal(oracle.sort_nodes(cnt, ulx), b)


def _geom(w, h):
    W, H = w - 32, h - 32
    nCols, nRows = int(np.float32(W) / np.float32(35)), int(np.float32(H) / np.float32(35))
    return W, H, nCols, int(math.ceil(np.float32(W) / nCols)), int(math.ceil(np.float32(H) / nRows))


@pytest.mark.parametrize("hh,ww,nf,seed,lt", [(480, 640, 1000, 2, False), (480, 640, 1000, 3, True),
                                              (720, 1280, 2000, 1, False), (480, 640, 5000, 6, False),
                                              (376, 1241, 2000, 8, False)])
def test_array_octree_equals_list_octree(oracle, hh, ww, nf, seed, lt):
    """octree_core.h (what the GPU CTA executes) vs DistributeOctTree restated
    with std::list/std::sort: same keypoints in the same order, independent of
    the order candidates arrive in."""
    lib = L.lib()
    ex = oracle.OracleExtractor(nf)
    ex.extract(synth_frame(hh, ww, seed, low_texture=lt))
    for lvl in range(8):
        w, h, q, _ = ex.level_info(lvl)
        W, H, nCols, wCell, hCell = _geom(w, h)
        c = ex.level_candidates(lvl)
        xys = np.stack([c["x"], c["y"], c["response"]], 1).astype(np.int32)
        ref = oracle.distribute(xys, W, H, q)
        xs = np.ascontiguousarray(xys[np.random.default_rng(lvl).permutation(len(xys))])
        out = np.zeros((q + 64, 3), np.int32)
        m = lib.orb_debug_octree_host(L.ptr(xs), len(xs), W, H, q, wCell, hCell, nCols, L.ptr(out), q + 64)
        assert m == len(ref) and np.array_equal(out[:m], ref), lvl


def test_array_octree_degenerate_inputs(oracle):
    lib = L.lib()
    out = np.zeros((64, 3), np.int32)
    empty = np.zeros((0, 3), np.int32)
    assert lib.orb_debug_octree_host(L.ptr(np.zeros((1, 3), np.int32)), 0, 608, 448, 10, 36, 38, 17,
                                     L.ptr(out), 64) == 0
    assert len(oracle.distribute(empty, 608, 448, 10)) == 0
    one = np.array([[100, 50, 33]], np.int32)
    assert lib.orb_debug_octree_host(L.ptr(one), 1, 608, 448, 10, 36, 38, 17, L.ptr(out), 64) == 1
    assert np.array_equal(out[:1], oracle.distribute(one, 608, 448, 10))
    # many points, tiny quota; equal responses everywhere (pure tie-breaking).  The
    # oracle breaks ties by list order, so feed it the canonical candidate order
    # (cell row, cell col, y, x) that ComputeKeyPointsOctTree produces.
    W, H, nCols, wCell, hCell = _geom(640, 480)
    rng = np.random.default_rng(1)
    pts = np.unique(rng.integers(3, 440, size=(3000, 2)), axis=0)
    pts[:, 0] += 100
    key = ((((pts[:, 1] - 3) // hCell) * nCols + (pts[:, 0] - 3) // wCell) << 14) | \
          (((pts[:, 1] - 3) % hCell) << 7) | ((pts[:, 0] - 3) % wCell)
    pts = pts[np.argsort(key)]
    xys = np.concatenate([pts, np.full((len(pts), 1), 40)], 1).astype(np.int32)
    shuffled = np.ascontiguousarray(xys[rng.permutation(len(xys))])
    for N in (1, 2, 5, 37, 300):
        ref = oracle.distribute(xys, W, H, N)
        outb = np.zeros((N + 64, 3), np.int32)
        m = lib.orb_debug_octree_host(L.ptr(shuffled), len(xys), W, H, N, wCell, hCell, nCols, L.ptr(outb), N + 64)
        assert m == len(ref) and np.array_equal(outb[:m], ref), N

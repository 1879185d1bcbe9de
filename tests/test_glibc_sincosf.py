"""csrc/glibc_sincosf.h == the libm of this box for EVERY float the extractor can pass (the reference's
`(float)cos(angle)` / `(float)sin(angle)`, ORBextractor.cc:111-112, resolve to cosf / sinf).  The host run is
exhaustive; the GPU test holds the device code against the host's libm on a dense sample + every angle that
distinguishes the formulations."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _libm():
    m = C.CDLL("libm.so.6")
    m.cosf.restype = m.sinf.restype = C.c_float
    m.cosf.argtypes = m.sinf.argtypes = [C.c_float]
    return m


def test_exhaustive_against_host_libm(tmp_path):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    exe = str(tmp_path / "sincosf_exhaustive")
    r = subprocess.run(["g++", "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-pthread",
                        os.path.join(ROOT, "tests", "native", "sincosf_exhaustive.cpp"), "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    out = subprocess.run([exe, str(min(os.cpu_count() or 1, 16))], capture_output=True, text=True, timeout=1200).stdout.split()
    n, bad_fused, bad_unfused, dbl, moves = (int(x) for x in out)
    assert n > 1_000_000_000
    assert bad_fused == 0 and bad_unfused == 0, out      # both builds of glibc's source agree with the restatement
    # the reason for the restatement: (float)cos((double)x) is NOT cosf -- if this ever becomes 0 the header is moot
    assert dbl > 0 and moves > 0, out


def test_host_hook_matches_libm_on_extractor_angles():
    from orb_slam3_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(0)
    deg = np.concatenate([rng.uniform(0, 360, 200000), np.arange(0, 360, 0.25)]).astype(np.float32)
    x = (deg * np.float32(np.pi / np.float32(180.0))).astype(np.float32)
    m = _libm()
    ref_c = np.array([m.cosf(float(v)) for v in x[:20000]], np.float32)
    ref_s = np.array([m.sinf(float(v)) for v in x[:20000]], np.float32)
    for fused in (1, 0):
        c, s = np.empty_like(x), np.empty_like(x)
        assert L.orb_debug_sincos_host(_lib.ptr(x), x.size, _lib.ptr(c), _lib.ptr(s), fused) == 0
        assert np.array_equal(c[:20000], ref_c) and np.array_equal(s[:20000], ref_s)


@pytest.mark.gpu
def test_device_sincos_equals_host_libm():
    """10^7 angles (uniform bit patterns in [0, 2*pi] + the extractor's angle lattice): device == host restatement
    (itself exhaustively equal to libm) == the oracle's libm call."""
    from orb_slam3_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(1)
    top = np.float32(6.2832).view(np.uint32)
    bits = rng.integers(0, int(top) + 1, 8_000_000, dtype=np.uint32)
    deg = rng.uniform(0, 360, 2_000_000).astype(np.float32)
    x = np.concatenate([bits.view(np.float32), (deg * np.float32(np.pi / np.float32(180.0))).astype(np.float32)])
    c, s, hc, hs = (np.empty_like(x) for _ in range(4))
    _lib.check(L.orb_debug_sincos_device(0, _lib.ptr(x), x.size, _lib.ptr(c), _lib.ptr(s)))
    assert L.orb_debug_sincos_host(_lib.ptr(x), x.size, _lib.ptr(hc), _lib.ptr(hs), 1) == 0
    assert np.array_equal(c, hc) and np.array_equal(s, hs)
    # and straight against libm through the oracle's helper on a slice
    m = _libm()
    idx = rng.integers(0, x.size, 50000)
    assert all(m.cosf(float(x[i])) == c[i] and m.sinf(float(x[i])) == s[i] for i in idx)

"""The DistributeOctTree kernel body (csrc/octree_core.h) run by several real host threads under ThreadSanitizer
(tests/native/octree_threads.cpp): the selection must equal the single-threaded run exactly and no pair of
conflicting accesses may be unordered.  Guards future changes of the latency-bound octree kernel on the CPU."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

from orb_slam3_b200.synth import synth_frame

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    exe = str(tmp_path_factory.mktemp("tsan") / "octree_threads")
    cmd = ["g++", "-std=c++20", "-O1", "-g", "-fsanitize=thread", "-pthread", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "native", "octree_threads.cpp"), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and "sanitize" in r.stderr:
        pytest.skip("ThreadSanitizer runtime not available: " + r.stderr[-200:])
    assert r.returncode == 0, r.stderr[-2000:]
    return exe


def _geom(w, h):
    W, H = w - 32, h - 32
    n_cols, n_rows = W // 35, H // 35
    return W, H, n_cols, int(np.ceil(W / n_cols)), int(np.ceil(H / n_rows))


@pytest.mark.parametrize("threads", [3, 8])
def test_octree_body_is_race_free_and_order_independent(oracle, harness, tmp_path, threads):
    ex = oracle.OracleExtractor(1000)
    ex.extract(synth_frame(480, 640, 9))
    for lvl in (0, 3, 7):
        w, h, q, _ = ex.level_info(lvl)
        W, H, n_cols, w_cell, h_cell = _geom(w, h)
        c = ex.level_candidates(lvl)
        xys = np.stack([c["x"], c["y"], c["response"]], 1).astype(np.int32)
        xys = np.ascontiguousarray(xys[np.random.default_rng(lvl).permutation(len(xys))])
        ref = oracle.distribute(xys, W, H, q)
        path = str(tmp_path / ("cand%d.bin" % lvl))
        with open(path, "wb") as f:
            f.write(struct.pack("<7i", len(xys), W, H, q, w_cell, h_cell, n_cols))
            f.write(xys.tobytes())
        env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 exitcode=66")
        r = subprocess.run([harness, path, str(threads)], capture_output=True, text=True, timeout=600, env=env)
        assert "ThreadSanitizer" not in r.stderr, r.stderr[:3000]
        assert r.returncode == 0, (r.returncode, r.stderr[:500])
        m_ref, m_thr, same = (int(x) for x in r.stdout.split())
        assert m_ref == m_thr == len(ref) and same == 1, (lvl, r.stdout)

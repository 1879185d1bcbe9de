"""The LocalInertialBA kernel body (csrc/lia_core.h) run by several real host threads under ThreadSanitizer
(tests/native/lia_threads.cpp): a pthread barrier plays __syncthreads(), CAS loops play the fp64 atomicAdd.
A clean run means every pair of conflicting accesses in the body is ordered by a barrier or is an atomic add --
the part of the device path a single-threaded host run cannot exercise -- and that the threaded result equals the
single-threaded one up to the summation order."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

from orb_slam3_b200 import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORDER = ["kf_Rwb", "kf_twb", "kf_Rcw", "kf_tcw", "kf_fixed", "kf_has_imu", "kf_vel", "kf_bg", "kf_ba", "mp_pos", "e_kf", "e_mp",
         "e_stereo", "e_obs", "e_inv_sigma2", "i_kf1", "i_kf2", "i_dR", "i_dV", "i_dP", "i_JRg", "i_JVg", "i_JVa", "i_JPg",
         "i_JPa", "i_bias", "i_dT", "i_C", "i_last"]


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    exe = str(tmp_path_factory.mktemp("tsan") / "lia_threads")
    cmd = ["g++", "-std=c++20", "-O1", "-g", "-fsanitize=thread", "-pthread", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "native", "lia_threads.cpp"), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and "sanitize" in r.stderr:
        pytest.skip("ThreadSanitizer runtime not available: " + r.stderr[-200:])
    assert r.returncode == 0, r.stderr[-2000:]
    return exe


def _dump(view, path):
    k = view._keep
    with open(path, "wb") as f:
        f.write(struct.pack("<5i", view.n_kf, view.n_mp, view.n_edges, view.n_inertial, view.iterations))
        f.write(struct.pack("<16d", *list(view.Rcb), *list(view.tcb), *list(view.tbc), view.lambda_init))
        f.write(struct.pack("<5f", view.fx, view.fy, view.cx, view.cy, view.bf))
        for name in ORDER:
            b = np.ascontiguousarray(k[name]).tobytes()
            f.write(struct.pack("<Q", len(b)))
            f.write(b)


@pytest.mark.parametrize("n_opt,n_mp,seed,perturb,threads", [(4, 80, 2, 1.0, 4), (6, 250, 1, 1.0, 8), (5, 120, 5, 6.0, 3)])
def test_kernel_body_is_race_free_under_real_threads(oracle, harness, tmp_path, n_opt, n_mp, seed, perturb, threads):
    d, _ = scenes.lia_scene(n_opt, n_mp, seed=seed, perturb=perturb)
    v = oracle.make_lia_view(d)
    path = str(tmp_path / "graph.bin")
    _dump(v, path)
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 exitcode=66")
    r = subprocess.run([harness, path, str(threads)], capture_output=True, text=True, timeout=600, env=env)
    assert "ThreadSanitizer" not in r.stderr, r.stderr[:3000]
    assert r.returncode == 0, (r.returncode, r.stderr[:1000])
    it_ref, it_thr, tr_ref, tr_thr, e_ref, e_thr, dpose, dpt, dchi = r.stdout.split()
    ref = oracle.lia_solve(v)
    assert int(it_ref) == ref["stats"]["iterations"] and int(tr_ref) == ref["stats"]["trials"]
    assert abs(int(it_thr) - int(it_ref)) <= 1
    if (it_thr, tr_thr) == (it_ref, tr_ref):
        assert float(dpose) < 1e-7 and float(dpt) < 1e-6 and float(dchi) < 1e-5, r.stdout
    assert abs(float(e_thr) - float(e_ref)) <= 1e-6 * float(e_ref)

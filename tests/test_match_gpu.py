"""GPU parity of the matchers (through the C ABI) against the CPU oracle: exact
assignment lists / pair lists and match counts on seeded scenes."""
import numpy as np
import pytest

from orb_slam3_b200 import scenes
from orb_slam3_b200.synth import synth_frame, shifted_frame

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def feats(oracle):
    out = {}
    for (h, w, nf, seed) in [(480, 640, 1000, 31), (720, 1280, 2000, 41)]:
        a = synth_frame(h, w, seed)
        b = shifted_frame(a, 5, -3, seed + 1)
        ex = oracle.OracleExtractor(nf)
        ka, da, _ = ex.extract(a)
        kb, db, _ = ex.extract(b)
        out[(h, w)] = (ka, da, kb, db)
    return out


@pytest.fixture(scope="module")
def matcher():
    from orb_slam3_b200.matcher import ORBmatcher
    return ORBmatcher


def test_descriptor_distance(oracle, matcher):
    rng = np.random.default_rng(1)
    for _ in range(100):
        a = rng.integers(0, 256, 32, dtype=np.uint8)
        b = rng.integers(0, 256, 32, dtype=np.uint8)
        assert matcher.DescriptorDistance(a, b) == oracle.ham_distance(a, b)


@pytest.mark.parametrize("size", [(480, 640), (720, 1280)])
@pytest.mark.parametrize("stereo", [False, True])
@pytest.mark.parametrize("th", [1.0, 3.0, 15.0])
def test_project_local_exact(oracle, matcher, feats, size, stereo, th):
    ka, da, _, _ = feats[size]
    F, mps = scenes.local_map_scene(ka, da, size[1], size[0], len(ka) // 2, seed=int(th) + 10 * stereo, stereo=stereo)
    for far in (False, True):
        n_ref, a_ref = oracle.match_project_local(F, mps, th, 0.8, far, 40.0)
        n, a = matcher(0.8).SearchByProjection(F, mps, th, far, 40.0)
        assert n == n_ref and np.array_equal(a, a_ref), (n, n_ref, int((a != a_ref).sum()))
    assert n_ref > 50


@pytest.mark.parametrize("size", [(480, 640), (720, 1280)])
@pytest.mark.parametrize("stereo", [False, True])
@pytest.mark.parametrize("seed", [2, 3])
def test_project_last_exact(oracle, matcher, feats, size, stereo, seed):
    ka, da, kb, db = feats[size]
    cur, last, Tcw = scenes.last_frame_scene(ka, da, kb, db, size[1], size[0], (5, -3), seed=seed, stereo=stereo)
    for th in (7.0, 15.0, 30.0):
        for (fw, bw) in ((0, 0), (1, 0), (0, 1)):
            for ori in (True, False):
                n_ref, a_ref = oracle.match_project_last(cur, last, Tcw, th, fw, bw, ori)
                n, a = matcher(0.9, ori).SearchByProjectionLast(cur, last, Tcw, th, fw, bw)
                assert n == n_ref and np.array_equal(a, a_ref), (th, fw, bw, ori, n, n_ref)
    assert n_ref > 50


def test_project_all_points_without_observations(oracle, matcher, feats):
    """Temporal points (Observations()==0) never block: heavy overwriting of slots."""
    ka, da, kb, db = feats[(480, 640)]
    cur, last, Tcw = scenes.last_frame_scene(ka, da, kb, db, 640, 480, (5, -3), seed=4, obs0_frac=1.0)
    n_ref, a_ref = oracle.match_project_last(cur, last, Tcw, 30.0, check_ori=True)
    n, a = matcher(0.9, True).SearchByProjectionLast(cur, last, Tcw, 30.0)
    assert n == n_ref and np.array_equal(a, a_ref)


@pytest.mark.parametrize("n_nodes", [40, 1000])
def test_triangulate_exact(oracle, matcher, feats, n_nodes):
    for size in ((480, 640), (720, 1280)):
        ka, da, kb, db = feats[size]
        k1, k2, fv1, fv2, F12, ep = scenes.triangulation_scene(ka, da, kb, db, size[1], size[0], seed=3,
                                                               n_nodes=n_nodes)
        for only_stereo in (False, True):
            for coarse in (False, True):
                for ori in (True, False):
                    n_ref, p_ref = oracle.match_triangulate(k1, k2, fv1, fv2, F12, ep, only_stereo, coarse, ori)
                    n, p = matcher(0.6, ori).SearchForTriangulation(k1, k2, fv1, fv2, F12, ep, only_stereo, coarse)
                    assert n == n_ref and np.array_equal(p, p_ref), (size, only_stereo, coarse, ori, n, n_ref)


def test_batches_equal_singles(oracle, matcher, feats):
    ka, da, kb, db = feats[(480, 640)]
    curs, lasts, Ts = [], [], []
    for s in range(5):
        c, l, T = scenes.last_frame_scene(ka, da, kb, db, 640, 480, (5, -3), seed=20 + s, stereo=bool(s % 2))
        curs.append(c); lasts.append(l); Ts.append(T)
    m = matcher(0.9, True)
    res, outs = m.project_last_batch(curs, lasts, np.stack(Ts), 15.0)
    for s in range(5):
        n_ref, a_ref = oracle.match_project_last(curs[s], lasts[s], Ts[s], 15.0)
        assert res[s] == n_ref and np.array_equal(outs[s], a_ref)
    Fs, Ms = zip(*[scenes.local_map_scene(ka, da, 640, 480, 500, seed=30 + s) for s in range(4)])
    res, outs = matcher(0.8).project_local_batch(list(Fs), list(Ms), 3.0)
    for s in range(4):
        n_ref, a_ref = oracle.match_project_local(Fs[s], Ms[s], 3.0, 0.8)
        assert res[s] == n_ref and np.array_equal(outs[s], a_ref)


def test_empty_and_degenerate(oracle, matcher, feats):
    ka, da, kb, db = feats[(480, 640)]
    F, mps = scenes.local_map_scene(ka, da, 640, 480, 10, seed=1)
    mps._keep["track_in_view"][:] = 0
    n, a = matcher(0.8).SearchByProjection(F, mps, 3.0)
    assert n == 0 and (a == -1).all()
    # identical descriptors everywhere: pure tie-breaking by candidate order
    d0 = np.zeros_like(da)
    F, mps = scenes.local_map_scene(ka, d0, 640, 480, 0, seed=2)
    mps._keep["desc"][:] = 0
    n_ref, a_ref = oracle.match_project_local(F, mps, 3.0, 0.8)
    n, a = matcher(0.8).SearchByProjection(F, mps, 3.0)
    assert n == n_ref and np.array_equal(a, a_ref)


def test_frame_keys_on_the_device_and_overflow_retry(oracle, feats):
    """on_device = 2: the frame's keypoints / descriptors are the extractor's device results, everything else is
    host memory -- same assignments as the all-host call.  With ORB_B200_MATCH_BUDGET=1 (subprocess: the budget is
    read when a matcher is created) every batch overflows its candidate buffer first and is re-run internally,
    synchronously and in asynchronous mode: the caller never sees it."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    code = r"""
import ctypes as C, numpy as np, sys, torch
sys.path.insert(0, %r); sys.path.insert(0, %r)
from orb_slam3_b200 import scenes
from orb_slam3_b200.extractor import ORBextractor
from orb_slam3_b200.matcher import ORBmatcher
from orb_slam3_b200.synth import synth_frame, shifted_frame
from orb_slam3_b200.views import orb_frame_view
from oracle import oracle
a = synth_frame(480, 640, 1); b = shifted_frame(a, 5, -3, 2)
ext = ORBextractor(1000, 1.2, 8, 20, 7)
_, ka, da = ext(a)
d = torch.from_numpy(np.stack([b, b])).cuda()
ext.extract_batch_device(d.data_ptr(), 2, 480, 640, 640, 480 * 640)
ext.synchronize()
_, kb, db = ext.download_results(0)
kp, ds, _, _, cap = ext.device_results()
curs, lasts, Ts, Fs, Ms = [], [], [], [], []
for s in range(2):
    c, l, T = scenes.last_frame_scene(ka, da, kb, db, 640, 480, (5, -3), seed=20 + s)
    F, M = scenes.local_map_scene(kb, db, 640, 480, 500, seed=30 + s)
    curs.append(c); lasts.append(l); Ts.append(T); Fs.append(F); Ms.append(M)
def on_dev(v, fr):
    w = orb_frame_view(); C.memmove(C.byref(w), C.byref(v), C.sizeof(w))
    w.keys = kp + fr * cap * 28; w.desc = ds + fr * cap * 32; w.u_right = None
    return w
m1, m2 = ORBmatcher(0.9, True), ORBmatcher(0.8, True)
r, outs = m1.project_last_batch([on_dev(c, i) for i, c in enumerate(curs)], lasts, np.stack(Ts), 15.0, on_device=2)
for s in range(2):
    n_ref, a_ref = oracle.match_project_last(curs[s], lasts[s], Ts[s], 15.0)
    assert r[s] == n_ref and np.array_equal(outs[s], a_ref), ("last", s)
r, outs = m2.project_local_batch([on_dev(f, i) for i, f in enumerate(Fs)], Ms, 3.0, on_device=2)
for s in range(2):
    n_ref, a_ref = oracle.match_project_local(Fs[s], Ms[s], 3.0, 0.8)
    assert r[s] == n_ref and np.array_equal(outs[s], a_ref), ("local", s)
# asynchronous, fully device-resident batch: overflow handled inside match_synchronize
m3 = ORBmatcher(0.8, True); m3.set_async(True)
dm = [{k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in M._keep.items()} for M in Ms]
from orb_slam3_b200.views import orb_mappoint_view
views, frames, keep = [], [], []
for i, (F, M) in enumerate(zip(Fs, Ms)):
    mv = orb_mappoint_view(); mv.n = M.n
    for k in ("track_in_view", "is_bad", "has_obs", "proj_x", "proj_y", "proj_xr", "scale_level", "view_cos", "depth", "desc"):
        setattr(mv, k, dm[i][k].data_ptr())
    fv = on_dev(F, i)
    tk = torch.from_numpy(F._keep[5]).cuda(); keep.append(tk)
    fv.kp_taken = tk.data_ptr()
    views.append(mv); frames.append(fv)
assign = torch.full((2, cap), -7, dtype=torch.int32, device="cuda")
r, _ = m3.project_local_batch(frames, views, 3.0, on_device=True, assign_ptrs=[assign.data_ptr() + 4 * i * cap for i in range(2)])
m3.synchronize()
for s in range(2):
    n_ref, a_ref = oracle.match_project_local(Fs[s], Ms[s], 3.0, 0.8)
    assert r[s] == n_ref and np.array_equal(assign[s, :len(a_ref)].cpu().numpy(), a_ref), ("async", s)
print("MATCH_MODES_OK")
""" % (os.path.dirname(here), here)
    for budget in ("48", "1"):
        env = dict(os.environ, ORB_B200_MATCH_BUDGET=budget)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0 and "MATCH_MODES_OK" in r.stdout, (budget, r.stdout[-1500:] + r.stderr[-3000:])

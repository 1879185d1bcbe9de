"""Randomised soak of the oracle against the reference's own object code (oracle/_ref): matchers,
isInFrustum, ComputeStereoMatches on random frame sizes / thresholds / poses, then the three optimisers with the reference's
LM driver plugged in.  Test infrastructure (imports oracle/); needs oracle/_ref built (`make -C oracle ref`).
Usage: python scripts/soak_ref.py   -> prints the number of comparisons and of mismatches (expected 0)."""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from orb_slam3_b200 import scenes
from orb_slam3_b200.synth import synth_frame, shifted_frame, stereo_right
from oracle import oracle as O, ref as R
bad = 0; total = 0
rng = np.random.default_rng(123)
for fs in range(6):
    h, w = [(480, 640), (720, 1280), (376, 1241)][fs % 3]
    a = synth_frame(h, w, 100 + fs); sh = (int(rng.integers(-8, 9)), int(rng.integers(-8, 9)))
    b = shifted_frame(a, sh[0], sh[1], 200 + fs)
    ex = O.OracleExtractor(int(rng.choice([500, 1000, 2000])))
    ka, da, _ = ex.extract(a); kb, db, _ = ex.extract(b)
    for seed in range(12):
        st = bool(seed % 2)
        th = float(rng.choice([1.0, 2.0, 3.0, 5.0, 15.0])); ratio = float(rng.choice([0.6, 0.75, 0.8, 0.9]))
        F, mps = scenes.local_map_scene(ka, da, w, h, int(rng.integers(0, 1500)), seed=seed + 50 * fs, stereo=st, th_noise=float(rng.uniform(0.3, 3)))
        far = bool(rng.integers(0, 2))
        r0 = O.match_project_local(F, mps, th, ratio, far, 35.0); r1 = R.front_project_local(F, mps, th, ratio, far, 35.0)
        ok = r0[0] == r1[0] and np.array_equal(r0[1], r1[1]); total += 1; bad += not ok
        if not ok: print("LOCAL mismatch", fs, seed, th, ratio, far)
        cur, last, Tcw = scenes.last_frame_scene(ka, da, kb, db, w, h, sh, seed=seed + 50 * fs, stereo=st, depth=float(rng.uniform(3, 20)))
        th2 = float(rng.choice([7.0, 15.0, 30.0])); fw, bw = [(0, 0), (1, 0), (0, 1)][int(rng.integers(0, 3))] if st else (0, 0); ori = bool(rng.integers(0, 2))
        r0 = O.match_project_last(cur, last, Tcw, th2, fw, bw, ori); r1 = R.front_project_last(cur, last, Tcw, th2, fw, bw, ori)
        ok = r0[0] == r1[0] and np.array_equal(np.where(r0[1] < 0, -1, r0[1]), r1[1]); total += 1; bad += not ok
        if not ok: print("LAST mismatch", fs, seed, th2, fw, bw, ori)
        k1, k2, fv1, fv2, _, _ = scenes.triangulation_scene(ka, da, kb, db, w, h, seed=seed + 50 * fs, stereo=st, n_nodes=int(rng.choice([40, 300, 1000])), shift=sh)
        T1 = np.concatenate([[0, 0, 0, 1], rng.normal(0, 0.3, 3)]).astype(np.float32)
        aa = rng.normal(0, 0.004, 3); thn = np.linalg.norm(aa)
        q = np.concatenate([aa / thn * np.sin(thn / 2), [np.cos(thn / 2)]])
        T2 = np.concatenate([q, T1[4:] + rng.normal(0, 0.08, 3)]).astype(np.float32)
        os_, co, ori = bool(rng.integers(0, 2)), bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
        n1, p1, F12, ep = R.front_triangulate(k1, k2, fv1, fv2, T1, T2, os_, co, ori)
        n0, p0 = O.match_triangulate(k1, k2, fv1, fv2, F12, ep, os_, co, ori)
        ok = n0 == n1 and np.array_equal(p0, p1); total += 1; bad += not ok
        if not ok: print("TRI mismatch", fs, seed, os_, co, ori, n0, n1)
    v, _ = scenes.frustum_scene(20000, seed=300 + fs)
    cl = float(rng.choice([0.0, 0.5, 0.9]))
    k0, o0 = O.is_in_frustum(v, cl); k1_, o1 = R.front_is_in_frustum(v, cl)
    ok = k0 == k1_ and all(np.array_equal(o0[k], o1[k]) for k in o0); total += 1; bad += not ok
    if not ok: print("FRUSTUM mismatch", fs)
    right = stereo_right(a, 400 + fs, disparities=tuple(int(x) for x in rng.integers(0, 60, int(rng.integers(1, 4)))), noise=int(rng.integers(0, 5)))
    el, er = O.OracleExtractor(1000), O.OracleExtractor(1000)
    kl, dl, _ = el.extract(a); kr, dr, _ = er.extract(right)
    pl = [el.level_image(l) for l in range(8)]; pr = [er.level_image(l) for l in range(8)]
    s0 = O.stereo_match(kl, dl, kr, dr, pl, pr, 386.0, 0.5514); s1 = R.front_stereo_match(kl, dl, kr, dr, pl, pr, 386.0, 0.5514)
    ok = s0[0] == s1[0] and np.array_equal(s0[1], s1[1]) and np.array_equal(s0[2], s1[2]); total += 1; bad += not ok
    if not ok: print("STEREO mismatch", fs)
print("front end: comparisons", total, "mismatches", bad)

# ---- the LM control law: restated vs the reference's driver, bit for bit
d = R.lm_driver()
lm_total = lm_bad = 0
for seed in range(40):
    if seed % 2:
        g, _ = scenes.lba_rough_graph(seed)
        lam = float(10.0 ** rng.uniform(-12, -4))
    else:
        g, _ = scenes.lba_graph(int(rng.integers(3, 15)), int(rng.integers(40, 800)), seed=seed, stereo_frac=float(rng.uniform(0, 1)))
        lam = 0.0
    gv = scenes.lba_view(g)
    iters = int(rng.integers(1, 25))
    a = O.lba_solve(gv, iters, lam); b = O.lba_solve(gv, iters, lam, driver=d)
    ok = (a["iterations"] == b["iterations"] and a["stats"]["trials"] == b["stats"]["trials"] and np.array_equal(a["kf_pose"], b["kf_pose"])
          and np.array_equal(a["mp_pos"], b["mp_pos"]) and np.array_equal(a["chi2"], b["chi2"]) and a["stats"]["lambda_final"] == b["stats"]["lambda_final"])
    lm_total += 1; lm_bad += not ok
    v, _ = scenes.pose_scene(int(rng.integers(3, 1200)), seed=seed, stereo_frac=float(rng.uniform(0, 1)), outlier_frac=float(rng.uniform(0, 0.4)))
    a = O.pose_optimize(v); b = O.pose_optimize(v, driver=d)
    ok = a["inliers"] == b["inliers"] and np.array_equal(a["pose"], b["pose"]) and np.array_equal(a["outlier"], b["outlier"]) and np.array_equal(a["stats"], b["stats"])
    lm_total += 1; lm_bad += not ok
    if seed < 12:
        dd, _ = scenes.lia_scene(int(rng.integers(3, 9)), int(rng.integers(40, 400)), seed=seed)
        lv = O.make_lia_view(dd)
        a = O.lia_solve(lv); b = O.lia_solve(lv, driver=d)
        ok = a["stats"] == b["stats"] and all(np.array_equal(a[k], b[k]) for k in ("Rcw", "tcw", "vel", "bg", "ba", "mp_pos", "chi2"))
        lm_total += 1; lm_bad += not ok
print("LM driver: comparisons", lm_total, "mismatches", lm_bad)

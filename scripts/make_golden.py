#!/usr/bin/env python3
"""Generate tests/golden/*.npz (run in the build container; needs cv2 + the oracle).

  primitives_cv2.npz : outputs of the un-vendored OpenCV primitives, computed by cv2 4.13 itself
                       (resize INTER_LINEAR, GaussianBlur 7x7 s2, FAST-9/16+NMS, fastAtan2) on
                       small seeded inputs -- the known answers the oracle is pinned to even
                       where cv2 is absent.
  extract_640x480.npz: oracle keypoints + descriptors of synth_frame(480,640,1), 1000 features
                       (BASELINE.json configs[0]).
  match_scene.npz    : oracle results of the three matchers on a seeded scene.
  lba_small.npz      : oracle optimize(10) result on lba_graph(8, 300, seed=1).
  lba_rig_small.npz  : the same on lba_rig_graph(8, 300, seed=1) (fisheye stereo rig, second-camera edges).
  stereo_640x480.npz : oracle Frame::ComputeStereoMatches on synth_frame(480,640,5) / stereo_right(.., 6, (5,30,17)).
  pose_small.npz     : oracle Optimizer::PoseOptimization on pose_scene(400, seed=7).
  frustum_small.npz  : oracle Frame::isInFrustum on frustum_scene(3000, seed=2).
  bow_small.npz      : oracle DBoW2 transform of the extract_640x480 descriptors on synth_vocabulary(10, 4, seed=2).
  lia_small.npz      : oracle Optimizer::LocalInertialBA optimize() on lia_scene(5, 150, seed=6).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cv2  # noqa: E402
from oracle import oracle as O  # noqa: E402
from orb_slam3_b200 import scenes  # noqa: E402
from orb_slam3_b200.synth import synth_frame, shifted_frame, stereo_right  # noqa: E402

out = os.path.join(ROOT, "tests", "golden")
os.makedirs(out, exist_ok=True)

# ---- cv2 primitives
rng = np.random.default_rng(7)
img = synth_frame(120, 160, 3)
rnd = rng.integers(0, 256, (97, 131), dtype=np.uint8)
det20 = cv2.FastFeatureDetector_create(threshold=20, nonmaxSuppression=True)
det7 = cv2.FastFeatureDetector_create(threshold=7, nonmaxSuppression=True)
fast = lambda d, im: np.array([(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in d.detect(im)], np.int32).reshape(-1, 3)
ys = rng.integers(-300000, 300000, 2000)
xs = rng.integers(-300000, 300000, 2000)
np.savez_compressed(
    os.path.join(out, "primitives_cv2.npz"), img=img, rnd=rnd,
    resize_img=cv2.resize(img, (133, 100), interpolation=cv2.INTER_LINEAR),
    resize_rnd=cv2.resize(rnd, (109, 81), interpolation=cv2.INTER_LINEAR),
    blur_img=cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101),
    blur_rnd=cv2.GaussianBlur(rnd, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101),
    fast20=fast(det20, img), fast7=fast(det7, img), atan_y=ys, atan_x=xs,
    atan=np.array([cv2.fastAtan2(float(y), float(x)) for y, x in zip(ys, xs)], np.float32),
    cv2_version=np.array(cv2.__version__))

# ---- extractor, config 0
f0 = synth_frame(480, 640, 1)
k, d, mono = O.OracleExtractor(1000).extract(f0)
np.savez_compressed(os.path.join(out, "extract_640x480.npz"), seed=1, kps=k, desc=d, mono=mono)

# ---- matchers
f1 = shifted_frame(f0, 5, -3, 2)
k1, d1, _ = O.OracleExtractor(1000).extract(f1)
cur, last, Tcw = scenes.last_frame_scene(k, d, k1, d1, 640, 480, (5, -3), seed=3, stereo=True)
n_last, a_last = O.match_project_last(cur, last, Tcw, 15.0)
F, mps = scenes.local_map_scene(k1, d1, 640, 480, 400, seed=4)
n_loc, a_loc = O.match_project_local(F, mps, 3.0, 0.8)
kf1, kf2, fv1, fv2, F12, ep = scenes.triangulation_scene(k, d, k1, d1, 640, 480, seed=5, n_nodes=60)
n_tri, pairs = O.match_triangulate(kf1, kf2, fv1, fv2, F12, ep)
np.savez_compressed(os.path.join(out, "match_scene.npz"), n_last=n_last, a_last=a_last, n_loc=n_loc, a_loc=a_loc,
                    n_tri=n_tri, pairs=pairs)

# ---- LBA
g, _ = scenes.lba_graph(8, 300, seed=1)
r = O.lba_solve(scenes.lba_view(g))
np.savez_compressed(os.path.join(out, "lba_small.npz"), kf_pose=r["kf_pose"], mp_pos=r["mp_pos"], chi2=r["chi2"],
                    iterations=r["iterations"], trials=r["stats"]["trials"], chi2_final=r["stats"]["chi2_final"])
# ---- LBA on a fisheye stereo rig: KannalaBrandt8 mono edges + EdgeSE3ProjectXYZToBody edges
g, _ = scenes.lba_rig_graph(8, 300, seed=1)
r = O.lba_solve(scenes.lba_view(g))
np.savez_compressed(os.path.join(out, "lba_rig_small.npz"), kf_pose=r["kf_pose"], mp_pos=r["mp_pos"], chi2=r["chi2"],
                    iterations=r["iterations"], trials=r["stats"]["trials"], chi2_final=r["stats"]["chi2_final"],
                    n_body=int((g["e_stereo"] == 2).sum()))
# ---- ComputeStereoMatches
sl = synth_frame(480, 640, 5)
sr = stereo_right(sl, 6, disparities=(5, 30, 17))
el, er = O.OracleExtractor(1000), O.OracleExtractor(1000)
kl, dl, _ = el.extract(sl)
kr, dr, _ = er.extract(sr)
n_st, ur, dp, sad = O.stereo_match(kl, dl, kr, dr, [el.level_image(l) for l in range(8)],
                                   [er.level_image(l) for l in range(8)], 386.0, 0.5514)
np.savez_compressed(os.path.join(out, "stereo_640x480.npz"), n=n_st, u_right=ur, depth=dp, sad=sad)

# ---- PoseOptimization
pv, _ = scenes.pose_scene(400, seed=7)
pr = O.pose_optimize(pv)
np.savez_compressed(os.path.join(out, "pose_small.npz"), inliers=pr["inliers"], pose=pr["pose"], outlier=pr["outlier"],
                    stats=pr["stats"])
# ---- isInFrustum
fv, _ = scenes.frustum_scene(3000, seed=2)
n_in, fo = O.is_in_frustum(fv, 0.5)
np.savez_compressed(os.path.join(out, "frustum_small.npz"), n_in=n_in, **fo)

# ---- ComputeBoW
voc = scenes.synth_vocabulary(10, 4, seed=2)
bw = O.bow_transform(voc, d, 2)
np.savez_compressed(os.path.join(out, "bow_small.npz"), **bw)

# ---- LocalInertialBA
ld, _ = scenes.lia_scene(5, 150, seed=6)
lr = O.lia_solve(O.make_lia_view(ld))
np.savez_compressed(os.path.join(out, "lia_small.npz"), tcw=lr["tcw"], Rcw=lr["Rcw"], vel=lr["vel"], bg=lr["bg"], ba=lr["ba"],
                    mp_pos=lr["mp_pos"], chi2=lr["chi2"], iterations=lr["stats"]["iterations"], trials=lr["stats"]["trials"],
                    err_end=lr["stats"]["err_end"])
print("golden fixtures written to", out)

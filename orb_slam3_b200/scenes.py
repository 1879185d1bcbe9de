"""Synthetic matching scenes on top of extracted keypoints (SURVEY.md 8d,
configs 2 and 3).  Pure numpy; shared by the tests and bench.py."""
import numpy as np

from .views import make_featvec_view, make_frame_view, make_lastframe_view, make_mappoint_view

FX = FY = 700.0


def scale_factors(nlevels=8, sf=1.2):
    s = [np.float32(1.0)]
    for _ in range(1, nlevels):
        s.append(np.float32(np.float64(s[-1]) * np.float64(np.float32(sf))))
    return np.array(s, np.float32)


def flip_bits(desc, nflip, rng):
    d = desc.copy()
    n = len(d)
    for _ in range(nflip):
        byte = rng.integers(0, 32, n)
        bit = rng.integers(0, 8, n)
        d[np.arange(n), byte] ^= (1 << bit).astype(np.uint8)
    return d


def stereo_u_right(kps, rng, frac=0.8, bf=386.0):
    """mvuRight: `frac` of the keypoints get a right coordinate from a random depth."""
    n = len(kps)
    depth = rng.uniform(4, 40, n).astype(np.float32)
    ur = (kps["x"] - np.float32(bf) / depth).astype(np.float32)
    ur[rng.random(n) > frac] = -1.0
    return ur


def last_frame_scene(last_kps, last_desc, cur_kps, cur_desc, width, height, shift, seed, stereo=False,
                     depth=8.0, obs0_frac=0.1, nomp_frac=0.15, taken_frac=0.05):
    """Frame t (last) -> frame t+1 (current) for SearchByProjection(Cur, Last): every
    last keypoint carries a MapPoint at `depth` in front of the last camera; the
    current pose is the pure translation that moves projections by `shift` px."""
    rng = np.random.default_rng(seed)
    sf = scale_factors()
    cx, cy = width / 2.0, height / 2.0
    n = len(last_kps)
    z = np.full(n, depth, np.float32) * rng.uniform(0.7, 1.4, n).astype(np.float32)
    X = np.stack([(last_kps["x"] - cx) / FX * z, (last_kps["y"] - cy) / FY * z, z], 1).astype(np.float32)
    t = np.array([shift[0] * depth / FX, shift[1] * depth / FY, 0.0], np.float32)
    Tcw = np.concatenate([[0, 0, 0, 1], t]).astype(np.float32)
    if seed % 2:  # a small rotation about z so the quaternion path is exercised
        a = 0.01
        Tcw[:4] = [0, 0, np.sin(a / 2), np.cos(a / 2)]
    has_mp = (rng.random(n) > nomp_frac)
    has_obs = (rng.random(n) > obs0_frac)
    mp_desc = flip_bits(last_desc, 3, rng)
    bf = 386.0 if stereo else 0.0
    ur = stereo_u_right(cur_kps, rng) if stereo else None
    taken = (rng.random(len(cur_kps)) < taken_frac).astype(np.uint8)
    cur = make_frame_view(cur_kps, cur_desc, width, height, sf, u_right=ur, kp_taken=taken, fx=FX, fy=FY,
                          cx=cx, cy=cy, bf=bf, b=bf / FX)
    last = make_lastframe_view(X, mp_desc, last_kps["octave"], last_kps["angle"], has_mp, has_obs)
    return cur, last, Tcw


def local_map_scene(kps, desc, width, height, n_extra, seed, stereo=False, th_noise=1.0, taken_frac=0.3):
    """SearchByProjection(F, vpMapPoints): one MapPoint per keypoint (projection =
    keypoint + N(0, th_noise) px, predicted level = octave) plus n_extra random ones."""
    rng = np.random.default_rng(seed)
    sf = scale_factors()
    n = len(kps)
    sel = rng.permutation(n)
    px = np.concatenate([kps["x"][sel] + rng.normal(0, th_noise, n), rng.uniform(20, width - 20, n_extra)])
    py = np.concatenate([kps["y"][sel] + rng.normal(0, th_noise, n), rng.uniform(20, height - 20, n_extra)])
    lvl = np.concatenate([kps["octave"][sel], rng.integers(0, 8, n_extra)])
    lvl = np.clip(lvl + (rng.random(n + n_extra) < 0.2) * rng.integers(-1, 2, n + n_extra), 0, 7)
    d = np.concatenate([flip_bits(desc[sel], 6, rng), rng.integers(0, 256, (n_extra, 32), dtype=np.uint8)])
    m = n + n_extra
    order = rng.permutation(m)
    vcos = np.where(rng.random(m) < 0.5, 0.9995, 0.9).astype(np.float32)
    depth = rng.uniform(2, 80, m).astype(np.float32)
    ur = stereo_u_right(kps, rng) if stereo else None
    pxr = (px - 386.0 / depth).astype(np.float32) if stereo else None
    mps = make_mappoint_view(px[order], py[order], lvl[order], d[order], view_cos=vcos[order],
                             proj_xr=None if pxr is None else pxr[order], depth=depth[order],
                             track_in_view=(rng.random(m) > 0.1), is_bad=(rng.random(m) < 0.02),
                             has_obs=(rng.random(m) > 0.1))
    taken = (rng.random(n) < taken_frac).astype(np.uint8)
    F = make_frame_view(kps, desc, width, height, sf, u_right=ur, kp_taken=taken, fx=FX, fy=FY, bf=386.0 if stereo else 0.0)
    return F, mps


def triangulation_scene(kps1, desc1, kps2, desc2, width, height, seed, n_nodes=1000, stereo=True,
                        shift=(5.0, -3.0)):
    """SearchForTriangulation: FeatureVectors bucket features by a descriptor hash
    (DBoW2 itself is out of scope), KF2 = KF1 translated along x (rectified pair
    geometry: F12 = [t]x for identical intrinsics up to scale)."""
    rng = np.random.default_rng(seed)
    sf = scale_factors()

    def nodes(desc):
        w = desc[:, :2].astype(np.int64)
        return ((w[:, 0] >> 3) * 32 + (w[:, 1] >> 3)) % n_nodes

    n1, n2 = len(kps1), len(kps2)
    node1, node2 = nodes(desc1), nodes(desc2)
    node1[rng.random(n1) < 0.02] = -1
    ur1 = stereo_u_right(kps1, rng, 0.6) if stereo else None
    ur2 = stereo_u_right(kps2, rng, 0.6) if stereo else None
    k1 = make_frame_view(kps1, desc1, width, height, sf, u_right=ur1, kp_taken=(rng.random(n1) < 0.4), fx=FX, fy=FY)
    k2 = make_frame_view(kps2, desc2, width, height, sf, u_right=ur2, kp_taken=(rng.random(n2) < 0.4), fx=FX, fy=FY)
    cx, cy = width / 2.0, height / 2.0
    K = np.array([[FX, 0, cx], [0, FY, cy], [0, 0, 1]], np.float64)
    R = np.eye(3)
    # translation (almost) parallel to the image plane along the image shift between the two
    # frames, so true correspondences satisfy the epipolar test
    t = np.array([shift[0] * 0.02, shift[1] * 0.02, 0.003])
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    F12 = (np.linalg.inv(K).T @ tx @ R @ np.linalg.inv(K)).astype(np.float32)
    ep = np.array([FX * t[0] / t[2] + cx, FY * t[1] / t[2] + cy], np.float32)
    return k1, k2, make_featvec_view(node1), make_featvec_view(node2), F12.reshape(9), ep

"""CPU pinning of the Frame::ComputeStereoMatches oracle (oracle/orc_stereo.cpp, Frame.cc:811-981):
an independent numpy restatement (cv2.norm for the window distance where cv2 is present), the
size-independent properties of a synthetic rectified pair, and the degenerate inputs."""
import numpy as np
import pytest

from orb_slam3_b200.synth import synth_frame, stereo_right

BF, B = 386.0, 0.5514


def _extract(oracle, left, right, nf):
    el, er = oracle.OracleExtractor(nf), oracle.OracleExtractor(nf)
    kl, dl, _ = el.extract(left)
    kr, dr, _ = er.extract(right)
    pl = [el.level_image(l) for l in range(8)]
    pr = [er.level_image(l) for l in range(8)]
    return kl, dl, kr, dr, pl, pr


def _numpy_restatement(kl, dl, kr, dr, pl, pr, bf, b):
    """Frame.cc:811-981 written independently of orc_stereo.cpp (python loops, numpy popcount)."""
    try:
        import cv2
        l1 = lambda a, c: float(cv2.norm(a, c, cv2.NORM_L1))
    except ImportError:  # the GPU box has no cv2
        l1 = lambda a, c: float(np.abs(a.astype(np.int32) - c.astype(np.int32)).sum())
    f32 = np.float32
    scale = np.ones(8, f32)
    for i in range(1, 8):
        scale[i] = f32(scale[i - 1] * f32(1.2))
    inv = (f32(1) / scale).astype(f32)
    N = len(kl)
    ur = np.full(N, -1, f32)
    dp = np.full(N, -1, f32)
    rows = [[] for _ in range(pl[0].shape[0])]
    for iR in range(len(kr)):
        r = f32(2) * scale[kr["octave"][iR]]
        for y in range(int(np.floor(kr["y"][iR] - r)), int(np.ceil(kr["y"][iR] + r)) + 1):
            rows[y].append(iR)
    maxD = f32(bf) / f32(b)
    lst = []
    bits = np.unpackbits(dr, axis=1)
    for iL in range(N):
        uL, vL, lv = kl["x"][iL], kl["y"][iL], kl["octave"][iL]
        cand = rows[int(vL)]
        if not cand or uL < 0:
            continue
        best, bi = 100, 0
        bl = np.unpackbits(dl[iL])
        for iR in cand:
            if abs(int(kr["octave"][iR]) - int(lv)) > 1:
                continue
            if f32(uL - maxD) <= kr["x"][iR] <= uL:
                d = int((bl != bits[iR]).sum())
                if d < best:
                    best, bi = d, iR
        if best >= 75:
            continue
        su, sv = np.round(f32(uL * inv[lv])), np.round(f32(vL * inv[lv]))
        # np.round is half-to-even; the products are never exactly .5 away from a grid point here
        sr = np.round(f32(kr["x"][bi] * inv[lv]))
        IL, IR = pl[lv], pr[lv]
        if sr < 0 or sr + 11 >= IL.shape[1]:
            continue
        y0, x0 = int(sv) - 5, int(su) - 5
        wl = np.ascontiguousarray(IL[y0:y0 + 11, x0:x0 + 11])
        d = [l1(wl, np.ascontiguousarray(IR[y0:y0 + 11, int(sr) + inc - 5:int(sr) + inc + 6])) for inc in range(-5, 6)]
        k = int(np.argmin(d))  # first minimum, like the strict '<' scan
        if k in (0, 10):
            continue
        d1, d2, d3 = f32(d[k - 1]), f32(d[k]), f32(d[k + 1])
        delta = f32(d1 - d3) / f32(f32(2) * f32(f32(d1 + d3) - f32(f32(2) * d2)))
        if delta < -1 or delta > 1:
            continue
        bu = f32(scale[lv] * f32(f32(f32(sr) + f32(k - 5)) + delta))
        disp = f32(uL - bu)
        if 0 <= disp < maxD:
            if disp <= 0:
                disp = f32(0.01)
                bu = f32(np.float64(uL) - 0.01)
            dp[iL] = f32(bf) / disp
            ur[iL] = bu
            lst.append((int(d[k]), iL))
    lst.sort()
    if lst:
        th = f32(f32(1.5) * f32(1.4)) * f32(lst[len(lst) // 2][0])
        for s, i in lst:
            if not f32(s) < th:
                ur[i] = dp[i] = -1
    return ur, dp


def test_oracle_matches_numpy_restatement(oracle):
    left = synth_frame(240, 320, 11)
    right = stereo_right(left, 12, disparities=(9, 21))
    args = _extract(oracle, left, right, 500)
    n, ur, dp, sad = oracle.stereo_match(*args, BF, B)
    ur2, dp2 = _numpy_restatement(*args, BF, B)
    assert n > 50 and n == int((ur >= 0).sum())
    assert np.array_equal(ur, ur2) and np.array_equal(dp, dp2)


@pytest.mark.parametrize("disp", [(12,), (5, 30, 17)])
def test_recovers_the_synthetic_disparity(oracle, disp):
    left = synth_frame(480, 640, 5)
    right = stereo_right(left, 6, disparities=disp)
    kl, dl, kr, dr, pl, pr = _extract(oracle, left, right, 1000)
    n, ur, dp, sad = oracle.stereo_match(kl, dl, kr, dr, pl, pr, BF, B)
    ok = ur >= 0
    assert n == ok.sum() and n > 300
    assert np.array_equal(ok, dp > 0)
    band = np.minimum((kl["y"] * len(disp) / 480).astype(int), len(disp) - 1)
    err = np.abs((kl["x"] - ur)[ok] - np.asarray(disp, np.float32)[band[ok]])
    assert np.median(err) < 0.3 and (err < 1.5).mean() > 0.9
    assert np.allclose(dp[ok], BF / (kl["x"] - ur)[ok], rtol=1e-6)
    # median gate (Frame.cc:968-982): everything kept lies below 2.1 x median of the candidates' distances
    s = np.sort(sad[sad >= 0])
    assert sad[ok].max() < np.float32(1.5) * np.float32(1.4) * np.float32(s[len(s) // 2])


def test_degenerate_inputs(oracle):
    left = synth_frame(240, 320, 3)
    kl, dl, kr, dr, pl, pr = _extract(oracle, left, left, 400)
    # no right keypoints at all: nothing matched, nothing undefined
    n, ur, dp, _ = oracle.stereo_match(kl, dl, kr[:0], dr[:0], pl, pr, BF, B)
    assert n == 0 and (ur == -1).all() and (dp == -1).all()
    # identical images: every window distance is 0, so is the median, and the '< 1.5*1.4*median' gate
    # (Frame.cc:970-981) rejects every match -- reference behaviour
    n, ur, dp, sad = oracle.stereo_match(kl, dl, kr, dr, pl, pr, BF, B)
    assert n == 0 and (ur == -1).all() and (sad >= 0).sum() > 50
    # zero true disparity + sensor noise: exact-zero disparities take the 'disparity <= 0 -> 0.01' branch
    # (Frame.cc:956-960), negative ones are dropped
    left = synth_frame(240, 320, 5)
    kl, dl, kr, dr, pl, pr = _extract(oracle, left, stereo_right(left, 105, disparities=(0,), noise=2), 400)
    n, ur, dp, _ = oracle.stereo_match(kl, dl, kr, dr, pl, pr, BF, B)
    ok = ur >= 0
    assert n > 100 and (kl["x"][ok] - ur[ok] > 0).all()
    zero = ok & (dp == np.float32(BF) / np.float32(0.01))
    assert zero.sum() >= 2
    assert np.array_equal(ur[zero], (kl["x"][zero].astype(np.float64) - 0.01).astype(np.float32))

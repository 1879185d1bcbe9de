"""Pin the CPU oracle's restatement of the un-vendored OpenCV primitives against
cv2 -- the reference itself ships no tests for this path.

cv2 is the ground truth for: resize(INTER_LINEAR), FAST-9/16 + NMS (values and
order), GaussianBlur 7x7 sigma 2, fastAtan2; and, composed, for every stage of
ORBextractor::operator() except the octree cull (reference-only logic)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from orb_slam3_b200.synth import synth_frame  # noqa: E402


def _cv_fast(img, th):
    det = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=True)
    return np.array([(int(p.pt[0]), int(p.pt[1]), int(p.response)) for p in det.detect(img)],
                    dtype=np.int32).reshape(-1, 3)


@pytest.mark.parametrize("shape,dst", [((720, 1280), (1067, 600)), ((480, 640), (533, 400)),
                                       ((201, 357), (298, 168)), ((97, 131), (109, 81))])
def test_resize_bit_exact(oracle, shape, dst):
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, size=shape, dtype=np.uint8)
    ref = cv2.resize(img, dst, interpolation=cv2.INTER_LINEAR)
    assert np.array_equal(ref, oracle.resize_linear_u8(img, dst[0], dst[1]))


def test_pyramid_chain_bit_exact(oracle):
    img = synth_frame(480, 640, 5)
    ex = oracle.OracleExtractor(1000)
    ex.extract(img)
    prev = img
    for lvl in range(1, 8):
        w, h, _, _ = ex.level_info(lvl)
        prev = cv2.resize(prev, (w, h), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(prev, ex.level_image(lvl)), lvl


@pytest.mark.parametrize("shape", [(200, 300), (33, 47), (7, 9), (64, 8)])
def test_blur_bit_exact(oracle, shape):
    rng = np.random.default_rng(4)
    img = rng.integers(0, 256, size=shape, dtype=np.uint8)
    ref = cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    assert np.array_equal(ref, oracle.gaussian_blur7(img))


@pytest.mark.parametrize("th", [7, 20, 40])
def test_fast_values_and_order(oracle, th):
    img = synth_frame(240, 320, 11)
    assert np.array_equal(_cv_fast(img, th), oracle.fast(img, th))
    cell = np.ascontiguousarray(img[50:91, 60:101])  # a 41x41 cell like ORBextractor.cc:805-872
    assert np.array_equal(_cv_fast(cell, th), oracle.fast(cell, th))


def test_fast_atan2(oracle):
    rng = np.random.default_rng(0)
    ys = rng.integers(-300000, 300000, 30000)
    xs = rng.integers(-300000, 300000, 30000)
    ys[:4] = [0, 0, 5, -5]
    xs[:4] = [0, 7, 0, 0]
    for y, x in zip(ys, xs):
        assert cv2.fastAtan2(float(y), float(x)) == oracle.fast_atan2(y, x)


def test_candidates_match_per_cell_cv2_calls(oracle):
    """ComputeKeyPointsOctTree's per-cell FAST with the minTh fallback (:805-872)."""
    fallbacks = 0
    for img in (synth_frame(480, 640, 2), synth_frame(480, 640, 3, low_texture=True)):
        ex = oracle.OracleExtractor(1000)
        ex.extract(img)
        for lvl in (0, 3, 7):
            im = ex.level_image(lvl)
            h, w = im.shape
            minB, maxBX, maxBY = 16, w - 16, h - 16
            width, height = float(maxBX - minB), float(maxBY - minB)
            nCols, nRows = int(width / 35), int(height / 35)
            wCell, hCell = int(np.ceil(width / nCols)), int(np.ceil(height / nRows))
            exp = []
            for i in range(nRows):
                iniY = minB + i * hCell
                maxY = min(iniY + hCell + 6, maxBY)
                if iniY >= maxBY - 3:
                    continue
                for j in range(nCols):
                    iniX = minB + j * wCell
                    maxX = min(iniX + wCell + 6, maxBX)
                    if iniX >= maxBX - 6:
                        continue
                    cell = np.ascontiguousarray(im[iniY:maxY, iniX:maxX])
                    k = _cv_fast(cell, 20)
                    if len(k) == 0:
                        k = _cv_fast(cell, 7)
                        fallbacks += 1
                    for x, y, r in k:
                        exp.append((x + j * wCell, y + i * hCell, r))
            got = ex.level_candidates(lvl)
            got = np.stack([got["x"], got["y"], got["response"]], 1).astype(np.int32)
            assert np.array_equal(np.array(exp, dtype=np.int32).reshape(-1, 3), got), lvl
    assert fallbacks > 0


def test_orientation_and_descriptor_vs_cv2_composition(oracle):
    """IC_Angle (:76-103) and computeOrbDescriptor (:107-146) recomposed from cv2 + numpy."""
    img = synth_frame(480, 640, 7)
    ex = oracle.OracleExtractor(1000)
    kps, desc, mono = ex.extract(img)
    assert mono == len(kps)
    umax = ex.umax()
    assert list(umax) == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    pat = oracle.pattern().reshape(512, 2)
    pos = 0
    for lvl in range(8):
        im = ex.level_image(lvl).astype(np.int64)
        blur = cv2.GaussianBlur(ex.level_image(lvl), (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
        lk = ex.level_keypoints(lvl)
        scale = ex.level_info(lvl)[3]
        for i in range(0, len(lk), 7):
            kp = lk[i]
            x, y = int(kp["x"]), int(kp["y"])
            m10 = m01 = 0
            for v in range(-15, 16):
                d = umax[abs(v)]
                row = im[y + v, x - d:x + d + 1]
                m10 += int((np.arange(-d, d + 1) * row).sum())
                m01 += v * int(row.sum())
            assert cv2.fastAtan2(float(m01), float(m10)) == kp["angle"]
            ang = np.float32(kp["angle"]) * np.float32(np.pi / np.float32(180.0))
            a, b = oracle.cos_sin_deg(kp["angle"])
            assert abs(a - np.cos(np.float64(ang))) < 1e-7 and abs(b - np.sin(np.float64(ang))) < 1e-7
            a, b = np.float32(a), np.float32(b)
            px = pat[:, 0].astype(np.float32)
            py = pat[:, 1].astype(np.float32)
            yy = np.rint(px * b + py * a).astype(np.int64)   # float32 products, float32 sum, half-even
            xx = np.rint(px * a - py * b).astype(np.int64)
            vals = blur[y + yy, x + xx].astype(np.int32)
            bits = (vals[0::2] < vals[1::2]).astype(np.uint8)
            exp = np.packbits(bits, bitorder="little")
            out = kps[pos + i]
            assert out["octave"] == lvl
            if lvl:
                assert out["x"] == np.float32(kp["x"]) * np.float32(scale)
            assert np.array_equal(exp, desc[pos + i]), (lvl, i)
        pos += len(lk)

"""GPU parity: liborbb200.so (through the C ABI / ORBextractor mirror) against
the CPU oracle on the same seeded frames -- bit-exact keypoints
(x, y, octave, angle, response, size), 32-byte descriptors, order and
monoIndex (BASELINE.json configs[0] and [1])."""
import numpy as np
import pytest

from orb_slam3_b200.synth import synth_frame, shifted_frame

pytestmark = pytest.mark.gpu


def _assert_same(ref, got, ctx=""):
    rk, rd, rm = ref
    gm, gk, gd = got
    assert gm == rm, (ctx, "monoIndex", gm, rm)
    assert len(gk) == len(rk), (ctx, "count", len(gk), len(rk))
    for f in ("x", "y", "size", "angle", "response", "octave", "class_id"):
        bad = np.nonzero(gk[f] != rk[f])[0]
        assert len(bad) == 0, (ctx, f, len(bad), bad[:5], gk[f][bad[:5]], rk[f][bad[:5]])
    assert np.array_equal(gd, rd), (ctx, "descriptors", int((gd != rd).any(1).sum()))


@pytest.fixture(scope="module")
def ext1000():
    from orb_slam3_b200.extractor import ORBextractor
    return ORBextractor(1000, 1.2, 8, 20, 7)


def test_pyramid_and_candidates_match_oracle(oracle, ext1000):
    img = synth_frame(480, 640, 2)
    ex = oracle.OracleExtractor(1000)
    ex.extract(img)
    ext1000(img)
    for lvl in range(8):
        assert np.array_equal(ex.level_image(lvl), ext1000.image_pyramid(lvl)), lvl
        c = ex.level_candidates(lvl)
        ref = np.stack([c["x"], c["y"], c["response"]], 1).astype(np.int32)
        got = ext1000.debug_candidates(lvl)
        key = lambda a: a[np.lexsort((a[:, 0], a[:, 1]))]
        assert len(ref) == len(got), (lvl, len(ref), len(got))
        assert np.array_equal(key(ref), key(got)), lvl


@pytest.mark.parametrize("seed,low", [(1, False), (2, False), (3, True), (4, False)])
def test_config1_640x480_1000_bit_exact(oracle, ext1000, seed, low):
    img = synth_frame(480, 640, seed, low_texture=low)
    ref = oracle.OracleExtractor(1000).extract(img)
    _assert_same(ref, ext1000(img), "seed%d" % seed)


def test_config2_1280x720_2000_bit_exact(oracle):
    from orb_slam3_b200.extractor import ORBextractor
    ext = ORBextractor(2000, 1.2, 8, 20, 7)
    orc = oracle.OracleExtractor(2000)
    for seed in (1, 7):
        img = synth_frame(720, 1280, seed)
        _assert_same(orc.extract(img), ext(img), "seed%d" % seed)


def test_lapping_area_and_mono_index(oracle, ext1000):
    img = synth_frame(480, 640, 5)
    for lap in ((0, 0), (0, 1000), (200, 400), (0, 320)):
        ref = oracle.OracleExtractor(1000).extract(img, lap)
        _assert_same(ref, ext1000(img, None, lap), str(lap))


def test_batch_equals_single(oracle):
    from orb_slam3_b200.extractor import ORBextractor
    ext = ORBextractor(1000, 1.2, 8, 20, 7)
    base = synth_frame(480, 640, 9)
    frames = [base]
    for t in range(1, 6):
        frames.append(shifted_frame(frames[-1], 3 - t, t - 2, 100 + t))
    res = ext.extract_batch(frames)
    orc = oracle.OracleExtractor(1000)
    for t, fr in enumerate(frames):
        _assert_same(orc.extract(fr), res[t], "frame%d" % t)


def test_other_geometries_and_parameters(oracle):
    from orb_slam3_b200.extractor import ORBextractor
    for (h, w, nf, sf, nl, ini, mn) in [(376, 1241, 2000, 1.2, 8, 20, 7), (480, 752, 1200, 1.2, 8, 20, 7),
                                        (400, 400, 500, 1.3, 5, 15, 5), (480, 640, 5000, 1.2, 8, 20, 7)]:
        img = synth_frame(h, w, 11)
        ref = oracle.OracleExtractor(nf, sf, nl, ini, mn).extract(img)
        got = ORBextractor(nf, sf, nl, ini, mn)(img)
        _assert_same(ref, got, str((h, w, nf)))


def test_strided_input_and_empty(oracle, ext1000):
    big = synth_frame(500, 700, 12)
    view = big[10:490, 30:670]  # 480x640 view with a 700-byte step
    ref = oracle.OracleExtractor(1000).extract(np.ascontiguousarray(view))
    _assert_same(ref, ext1000(view), "strided")
    mono, k, d = ext1000(np.zeros((0, 0), np.uint8))
    assert mono == -1 and len(k) == 0


def test_device_resident_path(oracle):
    import torch
    from orb_slam3_b200.extractor import ORBextractor
    from orb_slam3_b200._lib import KP_DTYPE
    ext = ORBextractor(1000, 1.2, 8, 20, 7)
    frames = np.stack([synth_frame(480, 640, 20 + i) for i in range(3)])
    d = torch.from_numpy(frames).cuda()
    ext.extract_batch_device(d.data_ptr(), 3, 480, 640, 640, 480 * 640)
    ext.synchronize()
    assert ext.device_results()[4] >= 1000
    orc = oracle.OracleExtractor(1000)
    for b in range(3):
        _assert_same(orc.extract(frames[b]), ext.download_results(b), "dev%d" % b)


def test_handles_with_different_parameters_interleave(oracle):
    """Kernel attributes (dynamic shared memory of the octree / resolve kernels) are per kernel, not per
    handle: a smaller handle used in between must not break a larger one."""
    from orb_slam3_b200.extractor import ORBextractor
    img = synth_frame(480, 640, 6)
    big, small = ORBextractor(2000, 1.2, 8, 20, 7), ORBextractor(300, 1.2, 8, 20, 7)
    ref_big = oracle.OracleExtractor(2000).extract(img)
    ref_small = oracle.OracleExtractor(300).extract(img)
    _assert_same(ref_big, big(img), "big first")
    _assert_same(ref_small, small(img), "small")
    _assert_same(ref_big, big(img), "big again")
    _assert_same(ref_small, small(img), "small again")


def test_large_device_batch_runs_in_lanes_and_results_are_double_buffered(oracle):
    """A device-resident batch of >= 32 frames is cut into sub-batches on separate streams (lanes); every frame
    must still equal the oracle.  The results of call i stay intact while call i+1 runs (double buffering)."""
    import torch
    from orb_slam3_b200.extractor import ORBextractor
    ext = ORBextractor(1000, 1.2, 8, 20, 7)
    uniq = [synth_frame(480, 640, 40 + i, low_texture=(i == 3)) for i in range(6)]
    frames = np.stack([uniq[i % 6] for i in range(40)])
    d = torch.from_numpy(frames).cuda()
    ext.extract_batch_device(d.data_ptr(), 40, 480, 640, 640, 480 * 640)
    ext.synchronize()
    kp0, ds0, n0, _, cap = ext.device_results()
    orc = oracle.OracleExtractor(1000)
    refs = [orc.extract(u) for u in uniq]
    for b in (0, 1, 3, 13, 14, 26, 27, 39):          # lane boundaries at 14 / 28 for 3 lanes
        _assert_same(refs[b % 6], ext.download_results(b), "lane frame %d" % b)
    # second call on other frames: the first call's device buffers must be untouched until the call after that
    d2 = torch.from_numpy(np.ascontiguousarray(frames[::-1])).cuda()
    ext.extract_batch_device(d2.data_ptr(), 40, 480, 640, 640, 480 * 640)
    ext.synchronize()
    kp1, ds1, _, _, _ = ext.device_results()
    assert kp1 != kp0 and ds1 != ds0                  # the other buffer set
    _assert_same(refs[(39 - 5) % 6], ext.download_results(5), "second call")
    ext.extract_batch_device(d.data_ptr(), 40, 480, 640, 640, 480 * 640)
    ext.synchronize()
    assert ext.device_results()[0] == kp0             # call i+2 reuses the first set

"""CPU pinning of the Frame::isInFrustum oracle (oracle/orc_frustum.cpp, Frame.cc:512-570 +
MapPoint.cc:505-546) and of the kernel body (csrc/frustum_core.h run on the host through
frustum_debug_host): an independent numpy float32 restatement with Eigen's a0 + (a1 + a2) sums,
exact agreement of the kernel source with the oracle including the stale-member semantics, and a
scene that takes every exit of the function."""
import numpy as np
import pytest

from orb_slam3_b200 import scenes

F32 = np.float32
KEYS = ("track_in_view", "proj_x", "proj_y", "proj_xr", "scale_level", "view_cos", "depth")


def _numpy_restatement(v, cos_limit):
    k = v._keep
    P, Pn = k["world_pos"], k["normal"]
    R = np.array(v.Rcw[:], F32).reshape(3, 3)
    t, Ow = np.array(v.tcw[:], F32), np.array(v.Ow[:], F32)
    s3 = lambda a, b, c: a + (b + c)
    Pc = np.stack([s3(R[r, 0] * P[:, 0], R[r, 1] * P[:, 1], R[r, 2] * P[:, 2]) + t[r] for r in range(3)], 1)
    dist_c = np.sqrt(s3(Pc[:, 0] * Pc[:, 0], Pc[:, 1] * Pc[:, 1], Pc[:, 2] * Pc[:, 2]))
    with np.errstate(divide="ignore", invalid="ignore"):
        invz = F32(1) / Pc[:, 2]
        u = F32(v.fx) * Pc[:, 0] / Pc[:, 2] + F32(v.cx)
        w = F32(v.fy) * Pc[:, 1] / Pc[:, 2] + F32(v.cy)
    front = ~(Pc[:, 2] < 0)
    inimg = front & ~((u < F32(v.min_x)) | (u > F32(v.max_x))) & ~((w < F32(v.min_y)) | (w > F32(v.max_y)))
    PO = P - Ow
    dist = np.sqrt(s3(PO[:, 0] * PO[:, 0], PO[:, 1] * PO[:, 1], PO[:, 2] * PO[:, 2]))
    band = inimg & ~((dist < F32(0.8) * k["min_dist"]) | (dist > F32(1.2) * k["max_dist"]))
    vc = s3(PO[:, 0] * Pn[:, 0], PO[:, 1] * Pn[:, 1], PO[:, 2] * Pn[:, 2]) / dist
    ok = band & ~(vc < F32(cos_limit))
    ratio = k["max_dist"] / dist
    q = np.log(ratio.astype(np.float64)) / np.float64(v.log_scale_factor)
    lvl = np.clip(np.ceil(q), 0, v.n_levels - 1).astype(np.int32)
    sure = np.abs(q - np.round(q)) > 1e-5        # float logf / division could land on the other side of an integer
    assert dist_c.dtype == F32 and vc.dtype == F32 and u.dtype == F32
    return dict(ok=ok, inimg=inimg, front=front, band=band, u=u, w=w, xr=u - F32(v.bf) * invz, vc=vc, depth=dist_c,
                lvl=lvl, sure=sure)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_oracle_matches_numpy_float32_restatement(oracle, seed):
    v, _ = scenes.frustum_scene(20000, seed=seed)
    n, o = oracle.is_in_frustum(v, 0.5)
    r = _numpy_restatement(v, 0.5)
    ok = r["ok"]
    assert n == ok.sum() and np.array_equal(o["track_in_view"].astype(bool), ok)
    # every exit of the function is taken by the scene
    assert (~r["front"]).sum() > 500 and (r["front"] & ~r["inimg"]).sum() > 500
    assert (r["inimg"] & ~r["band"]).sum() > 500 and (r["band"] & ~ok).sum() > 500 and ok.sum() > 500
    assert np.array_equal(o["proj_x"], np.where(r["inimg"], r["u"], F32(-1)))
    assert np.array_equal(o["proj_y"], np.where(r["inimg"], r["w"], F32(-1)))
    for key, ref in (("proj_xr", r["xr"]), ("view_cos", r["vc"]), ("depth", r["depth"])):
        assert np.array_equal(o[key][ok], ref[ok]), key
        assert (o[key][~ok] == 0).all()           # untouched (the wrapper starts from zeros)
    m = ok & r["sure"]
    assert np.array_equal(o["scale_level"][m], r["lvl"][m]) and m.sum() > 0.99 * ok.sum()
    assert set(np.unique(o["scale_level"][ok])) == set(range(8))


@pytest.mark.parametrize("cos_limit", [0.5, 0.0, 0.9])
def test_kernel_source_on_host_equals_oracle(oracle, cos_limit):
    from orb_slam3_b200 import frustum
    v, _ = scenes.frustum_scene(30000, seed=5)
    n_ref, ref = oracle.is_in_frustum(v, cos_limit)
    n, got = frustum.debug_host(v, cos_limit)
    assert n == n_ref
    for k in KEYS:
        assert np.array_equal(got[k], ref[k]), k


def test_stale_members_survive_a_second_frame(oracle):
    """Members the reference only writes for points in view keep the previous frame's values."""
    from orb_slam3_b200 import frustum
    v1, _ = scenes.frustum_scene(5000, seed=7)
    v2, _ = scenes.frustum_scene(5000, seed=8)
    _, ref = oracle.is_in_frustum(v1, 0.5)
    _, got = frustum.debug_host(v1, 0.5)
    n_ref, ref = oracle.is_in_frustum(v2, 0.5, out=ref)
    n, got = frustum.debug_host(v2, 0.5, out=got)
    assert n == n_ref
    for k in KEYS:
        assert np.array_equal(got[k], ref[k]), k
    gone = ref["track_in_view"] == 0
    assert (ref["depth"][gone] != 0).any()        # stale values from frame 1 are still there


def test_empty_and_bad_views():
    from orb_slam3_b200 import frustum
    from orb_slam3_b200._lib import OrbError
    v, _ = scenes.frustum_scene(0, seed=1)
    n, o = frustum.debug_host(v)
    assert n == 0 and len(o["proj_x"]) == 0
    v, _ = scenes.frustum_scene(10, seed=1)
    v.n_levels = 0
    with pytest.raises(OrbError):
        frustum.debug_host(v)

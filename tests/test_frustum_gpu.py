"""GPU parity of frame_is_in_frustum (C ABI, Frame::isInFrustum Frame.cc:512-570) against the CPU
oracle: every output bit-exact, including the members the reference leaves stale."""
import numpy as np
import pytest

from orb_slam3_b200 import scenes

pytestmark = pytest.mark.gpu

KEYS = ("track_in_view", "proj_x", "proj_y", "proj_xr", "scale_level", "view_cos", "depth")


@pytest.fixture(scope="module")
def culler():
    from orb_slam3_b200.frustum import FrustumCuller
    return FrustumCuller()


@pytest.mark.parametrize("n,seed,cos_limit", [(3000, 0, 0.5), (50000, 1, 0.5), (777, 2, 0.0), (20000, 3, 0.9)])
def test_all_outputs_bit_exact(oracle, culler, n, seed, cos_limit):
    v, _ = scenes.frustum_scene(n, seed=seed)
    n_ref, ref = oracle.is_in_frustum(v, cos_limit)
    n_got, got = culler.isInFrustum(v, cos_limit)
    assert n_got == n_ref and n_ref > n // 50
    for k in KEYS:
        bad = np.nonzero(got[k] != ref[k])[0]
        assert len(bad) == 0, (k, len(bad), bad[:5], got[k][bad[:5]], ref[k][bad[:5]])


def test_stale_members_and_handle_reuse(oracle, culler):
    v1, _ = scenes.frustum_scene(6000, seed=7)
    v2, _ = scenes.frustum_scene(4000, seed=8)
    v2b, _ = scenes.frustum_scene(6000, seed=9)
    _, ref = oracle.is_in_frustum(v1, 0.5)
    _, got = culler.isInFrustum(v1, 0.5)
    launches0 = culler.kernel_launches()
    culler.enqueue(v2, 0.5)                       # enqueue-only call in between: results stay on the device
    assert culler.device_results()["proj_x"] and culler.last_ms() > 0
    n_ref, ref = oracle.is_in_frustum(v2b, 0.5, out=ref)
    n_got, got = culler.isInFrustum(v2b, 0.5, out=got)
    assert n_got == n_ref and culler.kernel_launches() - launches0 == 2
    for k in KEYS:
        assert np.array_equal(got[k], ref[k]), k


def test_empty_and_bad_views(culler):
    from orb_slam3_b200._lib import OrbError
    v, _ = scenes.frustum_scene(0, seed=1)
    n, o = culler.isInFrustum(v)
    assert n == 0
    v, _ = scenes.frustum_scene(10, seed=1)
    v.log_scale_factor = 0.0
    with pytest.raises(OrbError):
        culler.isInFrustum(v)


def test_frustum_feeds_the_projection_matcher_on_the_device(oracle, culler):
    """Tracking::SearchLocalPoints end to end without the host in between: frame_is_in_frustum_device leaves
    the orb_mappoint_view fields on the device, match_project_local_batch(on_device) reads them there.
    Checked against the CPU chain oracle.is_in_frustum -> oracle.match_project_local."""
    import torch
    from orb_slam3_b200.matcher import ORBmatcher
    from orb_slam3_b200.synth import synth_frame
    from orb_slam3_b200.views import make_mappoint_view, orb_frame_view, orb_mappoint_view
    import ctypes as C
    kps, desc, _ = oracle.OracleExtractor(1000).extract(synth_frame(480, 640, 12))
    F, fv, mp_desc, is_bad, has_obs = scenes.frustum_match_scene(kps, desc, 640, 480, seed=4)
    # ---- CPU chain
    n_in, fo = oracle.is_in_frustum(fv, 0.5)
    mps = make_mappoint_view(fo["proj_x"], fo["proj_y"], fo["scale_level"], mp_desc, view_cos=fo["view_cos"],
                             proj_xr=fo["proj_xr"], depth=fo["depth"], track_in_view=fo["track_in_view"],
                             is_bad=is_bad, has_obs=has_obs)
    n_ref, a_ref = oracle.match_project_local(F, mps, 3.0, 0.8)
    assert n_in > 800 and n_ref > 300
    # ---- device chain on one stream
    st = torch.cuda.Stream()
    dev = {k: torch.from_numpy(np.ascontiguousarray(a)).cuda() for k, a in
           dict(keys=F._keep[0].view(np.uint8), fdesc=F._keep[1], taken=F._keep[5], mdesc=mp_desc, is_bad=is_bad,
                has_obs=has_obs).items()}
    d_assign = torch.empty(F.n, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    culler.enqueue(fv, 0.5, cuda_stream=st.cuda_stream)
    r = culler.device_results()
    Fd = orb_frame_view()
    C.memmove(C.byref(Fd), C.byref(F), C.sizeof(Fd))
    Fd.keys, Fd.desc, Fd.kp_taken, Fd.u_right = dev["keys"].data_ptr(), dev["fdesc"].data_ptr(), dev["taken"].data_ptr(), None
    Md = orb_mappoint_view()
    Md.n = fv.n
    for k in ("track_in_view", "proj_x", "proj_y", "proj_xr", "scale_level", "view_cos", "depth"):
        setattr(Md, k, r[k])
    Md.is_bad, Md.has_obs, Md.desc = dev["is_bad"].data_ptr(), dev["has_obs"].data_ptr(), dev["mdesc"].data_ptr()
    m = ORBmatcher(0.8)
    m.set_stream(st.cuda_stream)
    res, _ = m.project_local_batch([Fd], [Md], 3.0, on_device=True, assign_ptrs=[d_assign.data_ptr()])
    torch.cuda.synchronize()
    assert int(res[0]) == n_ref, (int(res[0]), n_ref)
    assert np.array_equal(d_assign.cpu().numpy(), a_ref)

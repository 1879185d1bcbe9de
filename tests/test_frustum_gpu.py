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

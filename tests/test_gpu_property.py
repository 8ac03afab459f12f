"""Property tests (hypothesis) on random shapes - SURVEY.md section 4 asks for them because the reference has no tests at all.

 * convolution: the tcgen05 kernels (TMA / cp.async producer, CTA pairs, stride 2, every epilogue variant) against the
   CUDA-core direct kernel on the same seeded random tensors - zero differing bytes;
 * scene: random frame sizes, depths, classes and constants against the CPU oracle - map / world / links bit-exact.
"""
import numpy as np
import pytest

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import HealthCheck, given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402

import oracle  # noqa: E402

pytestmark = pytest.mark.gpu

COMMON = dict(deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow], derandomize=True)


@settings(max_examples=40, **COMMON)
@given(tiles=st.integers(1, 9), h=st.integers(1, 30), w=st.integers(1, 30), ic16=st.integers(1, 20), oc=st.integers(1, 300),
       k=st.sampled_from([1, 3]), flags=st.sampled_from([0, 1, 2, 3, 4, 16]))
def test_conv_random_shapes(tod, tiles, h, w, ic16, oc, k, flags):
    from tod_b200 import _lib
    ms_tc, ms_direct, bad = _lib.conv_selftest(tiles, h, w, 16 * ic16, oc, k, iters=1, flags=flags)
    assert bad == 0


@settings(max_examples=15, **COMMON)
@given(tiles=st.integers(1, 40), hw=st.integers(2, 40), ic16=st.integers(1, 16), oc=st.integers(1, 260), flags=st.sampled_from([32, 33, 36]))
def test_conv_stride2_random_shapes(tod, tiles, hw, ic16, oc, flags):
    from tod_b200 import _lib
    ms_tc, ms_direct, bad = _lib.conv_selftest(tiles, hw, hw, 16 * ic16, oc, 3, iters=1, flags=flags)
    assert bad == 0


@settings(max_examples=10, **COMMON)
@given(tiles=st.integers(8, 40), hw=st.sampled_from([14, 20, 28]), ic128=st.integers(1, 3), oc32=st.integers(2, 8), flags=st.sampled_from([0, 1]))
def test_conv_pair_random_shapes(tod, tiles, hw, ic128, oc32, flags):
    """Shapes inside the CTA-pair kernel's range (>= 8 K iterations; the M-tile count decides per case)."""
    from tod_b200 import _lib
    ms_tc, ms_direct, bad = _lib.conv_selftest(tiles, hw, hw, 128 * ic128, 32 * oc32, 3, iters=1, flags=flags)
    assert bad == 0


@settings(max_examples=12, **COMMON)
@given(w=st.integers(34, 130), h=st.integers(24, 90), n=st.integers(1, 3), seed=st.integers(0, 10_000), weights_mode=st.sampled_from([0, 1]),
       sample_shift=st.sampled_from([0, 1]))
def test_scene_random_frames(tod, w, h, n, seed, weights_mode, sample_shift):
    rng = np.random.default_rng(seed)
    depth = rng.integers(0, 4001, (n, h, w)).astype(np.uint16)
    depth[rng.random((n, h, w)) < 0.1] = 0
    cls = rng.integers(0, 4, (n, h, w)).astype(np.uint16)
    ids = rng.integers(0, 100, (n, h, w)).astype(np.uint16)
    target = (cls | (ids << 8)).astype(np.uint16)
    params = dict(weights_mode=weights_mode, sample_shift=sample_shift)
    sb = tod.SceneBuilder(width=w, height=h, max_batch=n, **params)
    got = sb.append_batch(depth, target)
    op = oracle.scene_params(width=w, height=h, **params)
    for f in range(n):
        m, balls = oracle.pt_cloud(depth[f], target[f], op)
        world, c0, c1 = oracle.pt_cloud_weights(m, op)
        assert np.array_equal(got["map"][f], m)
        assert np.array_equal(got["world"][f].view(np.uint32), world.view(np.uint32))
        assert np.array_equal(got["conn1"][f].view(np.uint32), c1.view(np.uint32))
        assert np.array_equal(got["conn0"][f].view(np.uint32), c0.view(np.uint32))
        np.testing.assert_allclose(got["balls"][f], balls, rtol=1e-5, atol=0)

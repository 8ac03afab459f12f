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

import os  # noqa: E402

SCALE = max(1, int(os.environ.get("TOD_PROP_SCALE", "1")))  # TOD_PROP_SCALE=10: a ten times longer soak of the same properties
COMMON = dict(deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow], derandomize=True)


@settings(max_examples=40 * SCALE, **COMMON)
@given(tiles=st.integers(1, 9), h=st.integers(1, 30), w=st.integers(1, 30), ic16=st.integers(1, 20), oc=st.integers(1, 300),
       k=st.sampled_from([1, 3]), flags=st.sampled_from([0, 1, 2, 3, 4, 16]))
def test_conv_random_shapes(tod, tiles, h, w, ic16, oc, k, flags):
    from tod_b200 import _lib
    ms_tc, ms_direct, bad = _lib.conv_selftest(tiles, h, w, 16 * ic16, oc, k, iters=1, flags=flags)
    assert bad == 0


@settings(max_examples=15 * SCALE, **COMMON)
@given(tiles=st.integers(1, 40), hw=st.integers(2, 40), ic16=st.integers(1, 16), oc=st.integers(1, 260), flags=st.sampled_from([32, 33, 36]))
def test_conv_stride2_random_shapes(tod, tiles, hw, ic16, oc, flags):
    from tod_b200 import _lib
    ms_tc, ms_direct, bad = _lib.conv_selftest(tiles, hw, hw, 16 * ic16, oc, 3, iters=1, flags=flags)
    assert bad == 0


@settings(max_examples=10 * SCALE, **COMMON)
@given(tiles=st.integers(8, 40), hw=st.sampled_from([14, 20, 28]), ic128=st.integers(1, 3), oc32=st.integers(2, 8), flags=st.sampled_from([0, 1]))
def test_conv_pair_random_shapes(tod, tiles, hw, ic128, oc32, flags):
    """Shapes inside the CTA-pair kernel's range (>= 8 K iterations; the M-tile count decides per case)."""
    from tod_b200 import _lib
    ms_tc, ms_direct, bad = _lib.conv_selftest(tiles, hw, hw, 128 * ic128, 32 * oc32, 3, iters=1, flags=flags)
    assert bad == 0


@settings(max_examples=16 * SCALE, **COMMON)
@given(tiles=st.integers(40, 64), hw=st.integers(24, 60), ic16=st.integers(1, 16), oc=st.integers(1, 200), flags=st.sampled_from([0, 1, 2, 3]))
def test_conv_flat_many_tiles(tod, tiles, hw, ic16, oc, flags):
    """Dense 1x1 layers with more M tiles than SMs: the resident-weights path, the cp.async A producer (IC % 128 != 0), the
    linear staging + 1-D bulk store (16-byte rows) and the run-staged stores (other rows), with ragged last tiles."""
    from tod_b200 import _lib
    ms_tc, ms_direct, bad = _lib.conv_selftest(tiles, hw, hw, 16 * ic16, oc, 1, iters=1, flags=flags)
    assert bad == 0


@settings(max_examples=12 * SCALE, **COMMON)
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


def _micro_graph(S, c0, expand, dw_stride, oc, seed):
    """input -> QUANTIZE -> 3x3 s2 stem -> 1x1 expand -> (PAD +) depthwise -> 1x1 project -> RESIZE 2x -> 3x3 conv: the CUDA-core
    kernels (stem, depthwise, resize) and the tcgen05 convolutions around them on shapes the FRC graph does not have."""
    from oracle import synth_model as sm
    calib = sm.calib_images()[:, :S, :S]
    g = sm.Graph(seed, calib)
    x = g.input_u8([1, S, S, 3], float(sm.f32(1.0 / 128.0)), 128)
    x = g.quantize(x, float(sm.f32(1.0 / 128.0)), 0, sm.T_INT8, "input_int8")
    x = g.pad(x, ((0, 1), (0, 1)), "stem_pad")
    x = g.conv(x, c0, 3, 2, sm.PAD_VALID, sm.ACT_RELU6, "stem", gain=1.6)
    h = g.conv(x, c0 * expand, 1, 1, sm.PAD_SAME, sm.ACT_RELU6, "expand", gain=1.6)
    if dw_stride == 2:
        h = g.pad(h, ((0, 1), (0, 1)), "dw_pad")
        h = g.conv(h, 0, 3, 2, sm.PAD_VALID, sm.ACT_RELU6, "dw", gain=1.8, depthwise=True)
    else:
        h = g.conv(h, 0, 3, 1, sm.PAD_SAME, sm.ACT_RELU6, "dw", gain=1.8, depthwise=True)
    h = g.conv(h, oc, 1, 1, sm.PAD_SAME, sm.ACT_NONE, "project", gain=0.9)
    h = g.resize2x(h, "up")
    h = g.conv(h, oc, 3, 1, sm.PAD_SAME, sm.ACT_RELU, "smooth", gain=1.4)
    out = g.quantize(h, h.scale, h.zp + 128, sm.T_UINT8, "out")
    return g.serialize([g.tensors[0]], [out], "tod-b200 property-test micro graph")


@settings(max_examples=12 * SCALE, **COMMON)
@given(S=st.sampled_from([16, 24, 30, 40, 58]), c0=st.sampled_from([8, 16, 32]), expand=st.sampled_from([1, 3, 6]), dw_stride=st.sampled_from([1, 2]),
       oc=st.sampled_from([16, 32, 48]), seed=st.integers(1, 1000), fusion=st.sampled_from([0, 1]))
def test_micro_graph_every_tensor(tod, tmp_path_factory, S, c0, expand, dw_stride, oc, seed, fusion):
    blob = _micro_graph(S, c0, expand, dw_stride, oc, seed)
    path = tmp_path_factory.mktemp("micro") / "g.tflite"
    path.write_bytes(blob)
    from tests import synth
    tiles = synth.rgb_tiles(3, S=S, seed=seed)
    y = tod.Yolact.init(str(path), max_tiles=3, fusion=fusion, use_cuda_graph=fusion)
    res = y.infer_tiles(tiles, tile_classes=False, detections=False)
    for t in range(3):
        m = oracle.Model(str(path))
        m.invoke(tiles[t], threads=4)
        if fusion:
            want = m.tensor(m.outputs[0])
            assert np.array_equal(res["outputs"][0][t].reshape(-1), want.reshape(-1))
        else:
            for op in range(m.num_ops):
                ti = m.op_output(op)
                want = m.tensor(ti)
                got = y.fetch_tensor(ti, 3)[t]
                assert np.array_equal(got.reshape(-1), want.reshape(-1)), "tile %d op %d (code %d) differs in %d of %d" % (
                    t, op, m.op_code(op), (got.reshape(-1) != want.reshape(-1)).sum(), want.size)
    y.close()

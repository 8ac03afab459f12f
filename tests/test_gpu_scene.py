"""CUDA scene path vs the CPU oracle (pt_cloud.comp + pt_cloud_weights.comp restatement).

Bars: map u32 bit-exact; world/conn0/conn1 bit-exact (only IEEE mul/add/sqrt); balls <= 1e-5 relative.
All calls go through the C ABI (tod_scene_*).
"""
import numpy as np
import pytest

import oracle
from tests import synth

pytestmark = pytest.mark.gpu


def _compare(tod, depth, target, **params):
    n, H, W = depth.shape
    sb = tod.SceneBuilder(width=W, height=H, max_batch=n, **params)
    got = sb.append_batch(depth, target)
    op = oracle.scene_params(width=W, height=H, **params)
    for f in range(n):
        m, balls = oracle.pt_cloud(depth[f], target[f], op)
        world, c0, c1 = oracle.pt_cloud_weights(m, op)
        assert np.array_equal(got["map"][f], m), "map differs in frame %d (%d px)" % (f, (got["map"][f] != m).sum())
        assert np.array_equal(got["world"][f].view(np.uint32), world.view(np.uint32))
        assert np.array_equal(got["conn1"][f].view(np.uint32), c1.view(np.uint32))
        assert np.array_equal(got["conn0"][f].view(np.uint32), c0.view(np.uint32))
        np.testing.assert_allclose(got["balls"][f], balls, rtol=1e-5, atol=0)
    return sb, got


def test_terrain_only_literal(tod):
    depth = synth.depth_frames(2, seed=11)
    _compare(tod, depth, np.zeros_like(depth))  # target == 0 is what the reference feeds


def test_classes_and_intent_weights(tod):
    depth = synth.depth_frames(2, seed=12)
    target = synth.target_frames(2, seed=13)
    _compare(tod, depth, target, weights_mode=1)


def test_sample_shift_and_carmine_size(tod):
    depth = synth.depth_frames(3, W=320, H=240, seed=14)
    target = synth.target_frames(3, W=320, H=240, seed=15)
    _compare(tod, depth, target, sample_shift=1)


def test_edge_cases(tod):
    H, W = 480, 640
    depth = np.zeros((4, H, W), np.uint16)         # all holes: dz = 0 -> stamps hang over the bottom edge
    depth[1] = 65535                               # far beyond max depth: every stamp is off the top
    depth[2] = np.random.default_rng(0).integers(0, 65536, (H, W))  # noise: every pixel jumps rows
    depth[3] = 2000
    target = np.zeros_like(depth)
    target[3] = 1                                  # all robot: 40x40 stamps everywhere
    target[2, 100:200, 100:300] = 3 | (150 << 8)   # ball id >= 100 is ignored (unchecked in the shader)
    target[2, 300:310, 5:20] = 3 | (7 << 8)
    _compare(tod, depth, target)


def test_odd_size_and_constants(tod):
    depth = synth.depth_frames(1, W=100, H=75, seed=16)
    target = synth.target_frames(1, W=100, H=75, seed=17)
    _compare(tod, depth, target, terrain_norm_const=6, bot_norm_const=9, bot_avoidance_const=77.0, max_depth_in=3500.0)


def test_materialize_matches_scene_conversion(tod):
    depth = synth.depth_frames(2, seed=18)
    target = synth.target_frames(2, seed=19)
    sb, got = _compare(tod, depth, target)
    for f in range(2):
        s = sb.materialize(f)
        h, pos, b2, conn = oracle.scene_materialize(got["map"][f], got["world"][f], got["conn0"][f], got["conn1"][f], got["balls"][f])
        assert np.array_equal(s.height, h) and np.array_equal(s.pos, pos) and np.array_equal(s.connections, conn)
        assert np.array_equal(s.balls, b2)


def test_append_scene_pops_newest(tod):
    depth = synth.depth_frames(2, seed=20)
    dq, tq = [depth[0], depth[1]], [np.zeros_like(depth[0]), np.zeros_like(depth[1])]
    sb = tod.SceneBuilder(max_batch=1)
    scene = sb.append_scene(dq, tq)
    assert len(dq) == 1 and len(tq) == 1  # LIFO pop (scene.rs:186-187)
    m, _ = oracle.pt_cloud(depth[1], np.zeros_like(depth[1]))
    assert np.array_equal(scene.height, m.astype(np.float32).reshape(-1))


def test_batch_independence_and_capacity(tod):
    depth = synth.depth_frames(5, W=320, H=240, seed=21)
    target = synth.target_frames(5, W=320, H=240, seed=22)
    sb = tod.SceneBuilder(width=320, height=240, max_batch=5)
    a = sb.append_batch(depth, target)
    b = sb.append_batch(depth[::-1].copy(), target[::-1].copy())
    for k in ("map", "world", "conn0", "conn1", "balls"):
        assert np.array_equal(a[k], b[k][::-1]), k  # a frame's bytes do not depend on its position in the batch
    with pytest.raises(tod.TodError):
        tod.SceneBuilder(width=320, height=240, max_batch=2).append_batch(depth, target)


def test_stamp_kernels_agree(tod, monkeypatch):
    """TOD_STAMP_IMPL: 0 = pruned (default), 1 = shared-atomic, 2 = packed, 3 = generic - identical bytes, with robots and balls."""
    depth = synth.depth_frames(3, seed=31)
    target = synth.target_frames(3, seed=32, blobs=9)
    maps = []
    for impl in (0, 1, 2, 3):
        monkeypatch.setenv("TOD_STAMP_IMPL", str(impl))
        sb = tod.SceneBuilder(max_batch=3)
        maps.append(sb.append_batch(depth, target)["map"])
    m0, _ = oracle.pt_cloud(depth[0], target[0])
    assert np.array_equal(maps[0][0], m0)
    for k in (1, 2, 3):
        assert np.array_equal(maps[0], maps[k]), "stamp kernel %d differs from the pruned kernel" % k


def test_wide_terrain_bump_takes_the_generic_kernel(tod):
    """terrain_norm_const > 16 needs more than the 64 source columns the eight-warp atomic kernel stages (and a wider
    pattern than the pruned kernel's): the handle must fall back to the generic kernel and still match the oracle."""
    depth = synth.depth_frames(1, W=200, H=150, seed=33)
    target = synth.target_frames(1, W=200, H=150, seed=34)
    _compare(tod, depth, target, terrain_norm_const=17, bot_norm_const=20)
    _compare(tod, depth, target, terrain_norm_const=12, bot_norm_const=24)


def test_handles_of_different_sizes_coexist(tod, monkeypatch):
    """Opt-in shared memory is a per-function attribute: creating a smaller handle must not break an existing larger one."""
    depth = synth.depth_frames(1, seed=35)
    target = synth.target_frames(1, seed=36)
    small_d = synth.depth_frames(1, W=320, H=240, seed=37)
    for impl in ("1", "2", "3"):
        monkeypatch.setenv("TOD_STAMP_IMPL", impl)
        big = tod.SceneBuilder(max_batch=1)
        small = tod.SceneBuilder(width=320, height=240, max_batch=1)
        small.append_batch(small_d, np.zeros_like(small_d))
        got = big.append_batch(depth, target)
        m, _ = oracle.pt_cloud(depth[0], target[0])
        assert np.array_equal(got["map"][0], m)


def test_bench_size_batch_sampled_against_oracle(tod):
    """BASELINE config 3 at its size: 256 frames of 640x480 in one call; sampled frames against the oracle, and every
    frame against the same frame run alone (bytes do not depend on the batch)."""
    depth = synth.depth_frames(256, seed=3)
    target = synth.target_frames(256, seed=4)
    sb = tod.SceneBuilder(max_batch=256)
    got = sb.append_batch(depth, target, want=("map", "balls"))
    for f in (0, 1, 63, 128, 255):
        m, balls = oracle.pt_cloud(depth[f], target[f])
        assert np.array_equal(got["map"][f], m), "frame %d" % f
        np.testing.assert_allclose(got["balls"][f], balls, rtol=1e-5, atol=0)
    one = tod.SceneBuilder(max_batch=1)
    for f in (5, 77, 200):
        a = one.append_batch(depth[f:f + 1], target[f:f + 1])
        assert np.array_equal(a["map"][0], got["map"][f])
    full = sb.append_batch(depth[:8], target[:8])
    for f in (0, 7):
        world, c0, c1 = oracle.pt_cloud_weights(got["map"][f])
        assert np.array_equal(full["world"][f].view(np.uint32), world.view(np.uint32))
        assert np.array_equal(full["conn0"][f].view(np.uint32), c0.view(np.uint32))
        assert np.array_equal(full["conn1"][f].view(np.uint32), c1.view(np.uint32))


def test_map_only_device_call_and_materialize_guard(tod):
    """append_batch_device with no weight outputs takes the stand-alone tile merge; tod_scene_materialize refuses to run
    on results that went to caller buffers (it reads the handle's own images)."""
    torch = pytest.importorskip("torch")
    depth = synth.depth_frames(2, W=320, H=240, seed=41)
    target = synth.target_frames(2, W=320, H=240, seed=42)
    sb = tod.SceneBuilder(width=320, height=240, max_batch=2)
    want = sb.append_batch(depth, target)
    d_depth = torch.from_numpy(depth.view(np.int16)).cuda()
    d_target = torch.from_numpy(target.view(np.int16)).cuda()
    d_map = torch.zeros((2, 240, 320), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    sb.append_batch_device(d_depth.data_ptr(), d_target.data_ptr(), 2, d_map.data_ptr())
    with pytest.raises(tod.TodError):
        sb.materialize(0)
    torch.cuda.synchronize()
    sb.append_batch(depth[:1], target[:1], want=())   # drains the handle's stream
    assert np.array_equal(d_map.cpu().numpy().view(np.uint32), want["map"])
    # all outputs NULL: the results stay in the handle and materialize works
    sb.append_batch_device(d_depth.data_ptr(), d_target.data_ptr(), 2)
    s = sb.materialize(1)
    assert np.array_equal(s.height, want["map"][1].astype(np.float32).reshape(-1))

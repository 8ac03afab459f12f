"""BASELINE config 4: the fused RGB-D pipeline at Carmine resolution (320x240) with everything on the device -
classify (resize, two tiles, int8 graph, literal post-processing, resize back) writes the u16 target the point-cloud
kernels read, no host round trip (scene.rs:86-97 + scene.rs:147-331 in one stream).  Checked against the same path
composed through the host-facing calls, each of which is pinned against the oracle in test_gpu_post.py / test_gpu_scene.py."""
import numpy as np
import pytest

from tests import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("id_mode", [0, 1])
def test_fused_rgbd_device_pipeline(tod, models, id_mode):
    torch = pytest.importorskip("torch")
    full, _ = models
    W, H, n = 320, 240, 5
    frames = synth.rgb_frames(n, W=W, H=H, seed=51)
    depth = synth.depth_frames(n, W=W, H=H, seed=52)
    y = tod.Yolact.init(full, max_tiles=2 * n, id_mode=id_mode)
    sb = tod.SceneBuilder(width=W, height=H, max_batch=n)

    # host-composed path: classify in place, target = low 16 bits (scene.rs:93), then append
    host_frames = frames.copy()
    y.classify(host_frames, width=W, height=H)
    host_target = (host_frames & 0xFFFF).astype(np.uint16).reshape(n, H, W)
    want = sb.append_batch(depth, host_target)

    # device path: one stream, no host copies in between
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        d_frames = torch.from_numpy(frames.view(np.int32)).cuda()
        d_depth = torch.from_numpy(depth.view(np.int16)).cuda()
        d_target = torch.zeros((n, H, W), dtype=torch.int16, device="cuda")
        d_map = torch.zeros((n, H, W), dtype=torch.int32, device="cuda")
        d_world = torch.zeros((n, H, W, 4), dtype=torch.float32, device="cuda")
        d_c0, d_c1 = torch.zeros_like(d_world), torch.zeros_like(d_world)
        d_balls = torch.zeros((n, 100, 4), dtype=torch.float32, device="cuda")
        stream.synchronize()
        y.classify_device(d_frames.data_ptr(), n, W, H, d_target.data_ptr(), stream.cuda_stream)
        sb.append_batch_device(d_depth.data_ptr(), d_target.data_ptr(), n, d_map.data_ptr(), d_world.data_ptr(), d_c0.data_ptr(),
                               d_c1.data_ptr(), d_balls.data_ptr(), stream.cuda_stream)
        stream.synchronize()
    assert np.array_equal(d_frames.cpu().numpy().view(np.uint32), host_frames)
    assert np.array_equal(d_target.cpu().numpy().view(np.uint16), host_target)
    assert np.array_equal(d_map.cpu().numpy().view(np.uint32), want["map"])
    assert np.array_equal(d_world.cpu().numpy().view(np.uint32), want["world"].view(np.uint32))
    assert np.array_equal(d_c0.cpu().numpy().view(np.uint32), want["conn0"].view(np.uint32))
    assert np.array_equal(d_c1.cpu().numpy().view(np.uint32), want["conn1"].view(np.uint32))
    np.testing.assert_allclose(d_balls.cpu().numpy(), want["balls"], rtol=1e-5, atol=1e-6)

"""Parity at BASELINE.json's own sizes, on exactly the inputs bench.py times.

config 2: 64 tiles of synth.rgb_tiles(64, seed=2) through a max_tiles = 64 handle (the planner keys kernel choice on
          capacity: CTA pairs, resident weights and linear stores only appear at this size) - every output tensor, the
          literal class grid and the detections of ALL 64 tiles against the CPU oracle's committed CRCs
          (tests/golden/config2_oracle.json, tools/make_golden.py), and all 64 against the CUDA-core cross-check path;
          plus the reference's own two model-input fixtures (data/frc_balls.png, data/red_robot.png) as tiles.
config 4: 512 RGB-D frames at 320x240 through the fused device pipeline in one call, sampled frames against the oracle.
config 3 (256 depth frames) lives in tests/test_gpu_scene.py::test_bench_size_batch_sampled_against_oracle.
The frame sharder behind the C ABI (tod_pool_*) is checked against the single-handle calls on the same inputs.
"""
import hashlib
import json
import os
import zlib

import numpy as np
import pytest

import oracle
from tests import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes())


def golden(models):
    g = json.load(open(os.path.join(GOLD, "config2_oracle.json")))
    full, _ = models
    if hashlib.sha256(open(full, "rb").read()).hexdigest() != g["model_sha256"]:
        pytest.skip("the synthetic model on this machine differs from the one the oracle CRCs were made with (run tools/make_golden.py)")
    return g


def check_tile(rec, outputs, cells, det, t, what):
    for k in range(5):
        assert crc(outputs[k][t]) == rec["out"][k], "%s %d: output %d differs from the oracle" % (what, t, k)
    assert crc(cells[t]) == rec["cells"], "%s %d: class grid" % (what, t)
    assert int(det["count"][t]) == rec["n_det"], "%s %d: detection count" % (what, t)
    n = rec["n_det"]
    assert crc(det["priors"][t, :n]) == rec["prior"], "%s %d: NMS keep indices" % (what, t)
    assert crc(det["classes"][t, :n]) == rec["cls"], "%s %d: classes" % (what, t)
    assert crc(det["scores"][t, :n]) == rec["score"], "%s %d: scores" % (what, t)
    assert crc(det["boxes"][t, :n]) == rec["box"], "%s %d: boxes" % (what, t)


def raw_dets(y, n):
    """[n, max_dets, ...] detection arrays of the last call, as the C ABI lays them out (bit-packed masks only)"""
    det, keep = y._alloc_dets(n, True)
    det.masks = None
    det.masks_bin = None
    import ctypes as C
    from tod_b200._lib import check, lib
    check(lib().tod_yolact_fetch_detections(y._h, n, C.byref(det)))
    return keep


def test_config2_all_64_bench_tiles_against_oracle(tod, models):
    g = golden(models)
    full, _ = models
    tiles = synth.rgb_tiles(64, seed=2)
    y = tod.Yolact.init(full, max_tiles=64)
    res = y.infer_tiles(tiles, detections=True, cell_classes=True)
    keep = raw_dets(y, 64)
    assert not res["diverged"] or any(r["diverges"] for r in g["tiles"])
    for t in range(64):
        check_tile(g["tiles"][t], res["outputs"], res["cell_classes"], keep, t, "tile")
        # the replicated tile map is the grid x 8 (yolact.rs:127-128)
        assert np.array_equal(res["tile_classes"][t], np.kron(res["cell_classes"][t], np.ones((8, 8), np.uint32)))
    # float masks of a few tiles against the oracle (<= 1e-5 relative, IoU >= 0.999)
    o = y.outputs
    for t in (0, 31, 63):
        want = oracle.detect(res["outputs"][1][t], (o[1]["scale"], o[1]["zero_point"]), res["outputs"][0][t], (o[0]["scale"], o[0]["zero_point"]),
                             res["outputs"][2][t], (o[2]["scale"], o[2]["zero_point"]), res["outputs"][3][t], (o[3]["scale"], o[3]["zero_point"]))
        got = res["dets"][t]
        np.testing.assert_allclose(got["masks"], want["masks"], rtol=1e-5, atol=1e-7)
        inter, union = np.logical_and(got["masks_bin"], want["masks_bin"]).sum(), np.logical_or(got["masks_bin"], want["masks_bin"]).sum()
        assert union == 0 or inter / union >= 0.999
    # the whole batch once more through the CUDA-core direct convolutions (a third implementation of the same arithmetic)
    y.close()
    ref = tod.Yolact.init(full, max_tiles=64, conv_impl=1).infer_tiles(tiles, detections=False)
    for k in range(5):
        assert np.array_equal(ref["outputs"][k], res["outputs"][k]), "output %d: tcgen05 path differs from the direct path" % k


def test_reference_png_tiles(tod, models):
    """data/frc_balls.png and data/red_robot.png (SURVEY §8d config 1 substitute): the reference's only model inputs"""
    g = golden(models)
    full, _ = models
    tiles = np.load(os.path.join(GOLD, "ref_tiles.npz"))["tiles"]
    for mt in (2, 64):
        y = tod.Yolact.init(full, max_tiles=mt)
        res = y.infer_tiles(tiles, detections=True, cell_classes=True)
        keep = raw_dets(y, 2)
        for t in range(2):
            check_tile(g["ref_tiles"][t], res["outputs"], res["cell_classes"], keep, t, "png")
        y.close()


def test_pool_shards_tiles_like_one_handle(tod, models):
    """tod_pool_infer_tiles over 150 tiles (chunks of 64 over three handles per GPU, every visible GPU): the first 64 are
    the bench tiles (oracle CRCs), the rest against a single handle; bytes must not depend on handles or GPUs."""
    g = golden(models)
    full, _ = models
    tiles = np.concatenate([synth.rgb_tiles(64, seed=2), synth.rgb_tiles(86, seed=77)])
    y = tod.Yolact.init(full, max_tiles=64)
    pool = tod.Pool(full, depth=3, max_tiles=64)
    got = pool.infer_tiles(tiles, y.outputs, tile_classes=False, cell_classes=True, detections=True)
    for t in range(64):
        check_tile(g["tiles"][t], got["outputs"], got["cell_classes"], got["dets"], t, "pool tile")
    for lo in (64, 128):
        hi = min(150, lo + 64)
        want = y.infer_tiles(tiles[lo:hi], detections=True, float_masks=False, cell_classes=True)
        keep = raw_dets(y, hi - lo)
        for k in range(5):
            assert np.array_equal(got["outputs"][k][lo:hi], want["outputs"][k])
        assert np.array_equal(got["cell_classes"][lo:hi], want["cell_classes"])
        for key in ("count", "boxes", "scores", "classes", "priors", "masks_bits"):
            assert np.array_equal(got["dets"][key][lo:hi].view(np.uint32), keep[key].view(np.uint32)), key
    # a second pool on one GPU with one handle gives the same bytes (independent of G and depth)
    one = tod.Pool(full, devices=[0], depth=1, max_tiles=32)
    again = one.infer_tiles(tiles[:70], y.outputs, detections=True)
    for k in range(5):
        assert np.array_equal(again["outputs"][k], got["outputs"][k][:70])
    assert np.array_equal(again["dets"]["priors"], got["dets"]["priors"][:70])
    pool.close()
    one.close()


def test_pool_classify_matches_handle(tod, models):
    full, _ = models
    frames = synth.rgb_frames(9, seed=81)
    want = frames.copy()
    tod.Yolact.init(full, max_tiles=18).classify(want)
    pool = tod.Pool(full, depth=2, max_tiles=4)   # two frames per chunk
    got = frames.copy()
    pool.classify(got)
    assert np.array_equal(got, want)
    pool.close()


def _oracle_classify(m, frame, W, H):
    seg_t = m.outputs[4]
    info = m.tensor_info(seg_t)
    tiles = oracle.classify_pre(frame, W, H)
    outs = []
    for t in range(2):
        m.invoke(tiles[t], threads=8)
        px, _ = oracle.postprocess_tile(m.tensor(seg_t), info["scale"], info["zero_point"], mode=0)
        outs.append(px)
    return oracle.classify_post(outs[0], outs[1], W, H)


def test_config4_512_fused_frames_sampled_against_oracle(tod, models):
    """BASELINE config 4: 512 RGB-D frames, 320x240, classify -> target -> point cloud + weights on the device in one call."""
    torch = pytest.importorskip("torch")
    full, _ = models
    W, H, n = 320, 240, 512
    frames = np.tile(synth.rgb_frames(16, W=W, H=H, seed=5), (n // 16, 1))
    depth = np.tile(synth.depth_frames(16, W=W, H=H, seed=3), (n // 16, 1, 1))
    frames[300:] ^= np.uint32(0x10203000)   # make the second half differ from the first
    y = tod.Yolact.init(full, max_tiles=2 * n)
    sb = tod.SceneBuilder(width=W, height=H, max_batch=n)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        d_frames = torch.from_numpy(frames.view(np.int32)).cuda()
        d_depth = torch.from_numpy(depth.view(np.int16)).cuda()
        d_target = torch.zeros((n, H, W), dtype=torch.int16, device="cuda")
        d_map = torch.zeros((n, H, W), dtype=torch.int32, device="cuda")
        d_world = torch.zeros((n, H, W, 4), dtype=torch.float32, device="cuda")
        d_c0, d_c1 = torch.zeros_like(d_world), torch.zeros_like(d_world)
        stream.synchronize()
        y.classify_device(d_frames.data_ptr(), n, W, H, d_target.data_ptr(), stream.cuda_stream)
        sb.append_batch_device(d_depth.data_ptr(), d_target.data_ptr(), n, d_map.data_ptr(), d_world.data_ptr(), d_c0.data_ptr(), d_c1.data_ptr(), None,
                               stream.cuda_stream)
        stream.synchronize()
    out_frames = d_frames.cpu().numpy().view(np.uint32)
    out_map = d_map.cpu().numpy().view(np.uint32)
    m = oracle.Model(full)
    op = oracle.scene_params(width=W, height=H)
    for f in (0, 257, 511):
        wf = _oracle_classify(m, frames[f], W, H)
        assert np.array_equal(out_frames[f], wf), "frame %d: %d px differ" % (f, (out_frames[f] != wf).sum())
        target = oracle.target_from_frame(wf).reshape(H, W)
        wm, _ = oracle.pt_cloud(depth[f], target, op)
        assert np.array_equal(out_map[f], wm), "frame %d: map" % f
        world, c0, c1 = oracle.pt_cloud_weights(wm, op)
        assert np.array_equal(d_world[f].cpu().numpy().view(np.uint32), world.view(np.uint32))
        assert np.array_equal(d_c0[f].cpu().numpy().view(np.uint32), c0.view(np.uint32))
        assert np.array_equal(d_c1[f].cpu().numpy().view(np.uint32), c1.view(np.uint32))
    # identical input frames give identical bytes wherever they sit in the batch
    assert np.array_equal(out_frames[0], out_frames[16]) and np.array_equal(out_map[3], out_map[19])
    # the in-library RGB-D loop (tod_pool_rgbd_batch: chunks of 32 frames over the pool's handles) gives the same bytes
    del d_world, d_c0, d_c1
    y.close()
    pool = tod.Pool(full, depth=2, max_tiles=64)
    pf = frames[:96].copy()
    got = pool.rgbd(pf, depth[:96], sb.params, want=("map", "conn0"))
    assert np.array_equal(pf, out_frames[:96])
    assert np.array_equal(got["map"], out_map[:96])
    pool.close()

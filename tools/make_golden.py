#!/usr/bin/env python
"""Regenerates the committed fixtures under tests/golden/ (run in the build container, where /root/reference exists).

  bmp_trio.npz      the reference's own debug dumps (/root/reference/depth.bmp, image.bmp, map.bmp; written by
                    src/scene.rs:192-194, 306-309, 291-293 during a dev run): the recovered u16 depth frame, the mask
                    of pixels whose depth could be recovered exactly, and the low byte of the reference's map.
  oracle_golden.npz outputs of the CPU oracle on seeded inputs (CRC32 per array + a few small arrays), so that both a
                    change of the oracle and a drift of the CUDA path are caught against a committed value.
  ref_tiles.npz     the reference's only model-input fixtures, /root/reference/data/frc_balls.png and red_robot.png
                    (224x224 RGB = exactly one tile, data/README.md:5-14), decoded to u8[2,224,224,3].
  config2_oracle.json  BASELINE config 2 at its size: CRC32 of every output tensor, of the literal class grid and of the
                    detection keep indices, per tile, for the 64 bench tiles (synth.rgb_tiles(64, seed=2)) and the two
                    reference PNG tiles, computed by the CPU oracle.  bench.py and tests/test_gpu_baseline_size.py compare
                    the CUDA path's bytes with these.
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle import synth_model  # noqa: E402
from tests import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def bmp_trio():
    import cv2
    ref = "/root/reference"
    d17 = cv2.imread(ref + "/depth.bmp", cv2.IMREAD_UNCHANGED)   # depth / 17 as u8
    img = cv2.imread(ref + "/image.bmp", cv2.IMREAD_UNCHANGED)   # sampled depth & 0xFF (dbg.comp), shifted by (+1,+1)
    mp = cv2.imread(ref + "/map.bmp", cv2.IMREAD_UNCHANGED)      # map as u8
    lo = np.roll(img, (-1, -1), axis=(0, 1)).astype(np.int32)
    q = d17.astype(np.int32)
    depth = 17 * q + 8
    ok = np.zeros(q.shape, bool)
    for r in range(17):
        dd = 17 * q + r
        m = ((dd & 0xFF) == lo) & ~ok
        depth[m] = dd[m]
        ok |= m
    np.savez_compressed(os.path.join(OUT, "bmp_trio.npz"), depth=depth.astype(np.uint16), exact=np.packbits(ok), map_lo8=mp)
    print("bmp_trio: recovered %.1f%% of the depth pixels exactly" % (100 * ok.mean()))


def oracle_golden():
    full, small = synth_model.ensure_models()
    g = {}
    # scene path
    depth = synth.depth_frames(2, W=320, H=240, seed=101)
    target = synth.target_frames(2, W=320, H=240, seed=102)
    for mode in (0, 1):
        p = oracle.scene_params(width=320, height=240, weights_mode=mode)
        for f in range(2):
            m, balls = oracle.pt_cloud(depth[f], target[f], p)
            w, c0, c1 = oracle.pt_cloud_weights(m, p)
            g["scene_m%d_f%d" % (mode, f)] = np.array([crc(m), crc(w), crc(c0), crc(c1)], np.uint32)
            g["scene_balls_m%d_f%d" % (mode, f)] = balls
    # graph: every output of both models on seeded tiles
    for name, path, S in (("small", small, 64), ("full", full, 224)):
        tiles = synth.rgb_tiles(2, S=S, seed=103)
        m = oracle.Model(path)
        for t in range(2):
            m.invoke(tiles[t], threads=8)
            g["%s_out_t%d" % (name, t)] = np.array([crc(m.tensor(o)) for o in m.outputs], np.uint32)
            if name == "small":
                g["small_all_t%d" % t] = np.array([crc(m.tensor(m.op_output(i))) for i in range(m.num_ops)], np.uint32)
            if name == "full":
                outs = [m.tensor(o) for o in m.outputs]
                info = [m.tensor_info(o) for o in m.outputs]
                qp = lambda i: (info[i]["scale"], info[i]["zero_point"])
                d = oracle.detect(outs[1], qp(1), outs[0], qp(0), outs[2], qp(2), outs[3], qp(3))
                g["full_det_prior_t%d" % t] = d["prior"].astype(np.int32)
                g["full_det_cls_t%d" % t] = d["cls"].astype(np.int32)
                g["full_det_score_t%d" % t] = d["score"]
                px, div = oracle.postprocess_tile(outs[4], info[4]["scale"], info[4]["zero_point"], 0)
                g["full_cls_t%d" % t] = np.array([crc(px), int(div)], np.uint32)
    # classify() pipeline on one frame
    frame = synth.rgb_frames(1, seed=104)[0]
    tiles = oracle.classify_pre(frame)
    g["classify_pre"] = np.array([crc(tiles)], np.uint32)
    np.savez_compressed(os.path.join(OUT, "oracle_golden.npz"), **g)
    print("oracle_golden: %d entries" % len(g))


def ref_tiles():
    import cv2
    out = []
    for name in ("frc_balls.png", "red_robot.png"):
        bgr = cv2.imread("/root/reference/data/" + name, cv2.IMREAD_COLOR)
        assert bgr.shape == (224, 224, 3), bgr.shape
        out.append(bgr[..., ::-1].copy())
    np.savez_compressed(os.path.join(OUT, "ref_tiles.npz"), tiles=np.stack(out).astype(np.uint8), names=np.array(["frc_balls.png", "red_robot.png"]))
    print("ref_tiles: 2 tiles")


def tile_record(m, tile, threads=8):
    """per-tile CRCs the CUDA path is compared with: the five outputs, the literal 28x28 class grid, detections"""
    m.invoke(tile, threads=threads)
    outs = [m.tensor(o) for o in m.outputs]
    info = [m.tensor_info(o) for o in m.outputs]
    qp = lambda i: (info[i]["scale"], info[i]["zero_point"])
    d = oracle.detect(outs[1], qp(1), outs[0], qp(0), outs[2], qp(2), outs[3], qp(3))
    px, div = oracle.postprocess_tile(outs[4], info[4]["scale"], info[4]["zero_point"], 0)
    cells = np.ascontiguousarray(px.reshape(224, 224)[::8, ::8])
    return {"out": [int(crc(o)) for o in outs], "cells": int(crc(cells)), "diverges": bool(div), "n_det": int(d["n"]),
            "prior": int(crc(d["prior"].astype(np.int32))), "cls": int(crc(d["cls"].astype(np.int32))), "score": int(crc(d["score"].astype(np.float32))),
            "box": int(crc(d["box"].astype(np.float32)))}


def config2_oracle():
    import hashlib
    import json
    full, _ = synth_model.ensure_models()
    m = oracle.Model(full)
    tiles = synth.rgb_tiles(64, seed=2)
    ref = np.load(os.path.join(OUT, "ref_tiles.npz"))["tiles"]
    rec = {"model_sha256": hashlib.sha256(open(full, "rb").read()).hexdigest(), "input": "tests.synth.rgb_tiles(64, seed=2)",
           "tiles": [tile_record(m, tiles[t]) for t in range(64)], "ref_tiles": [tile_record(m, ref[t]) for t in range(2)]}
    json.dump(rec, open(os.path.join(OUT, "config2_oracle.json"), "w"), indent=0)
    print("config2_oracle: %d + %d tiles, detections per tile %s" % (len(rec["tiles"]), len(rec["ref_tiles"]), sorted({r["n_det"] for r in rec["tiles"]})))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["bmp_trio", "oracle_golden", "ref_tiles", "config2_oracle"]
    for w in which:
        globals()[w]()

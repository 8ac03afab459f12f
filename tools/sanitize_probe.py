#!/usr/bin/env python
"""Small end-to-end run of every new kernel family, for compute-sanitizer (memcheck / racecheck / initcheck):
scene path at two sizes with robots and balls (pruned stamp + merge + both weights kernels), the path relaxation,
the small YOLACT model with detections.    compute-sanitizer --tool racecheck python tools/sanitize_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import tod_b200  # noqa: E402
from oracle import synth_model  # noqa: E402
from tests import synth  # noqa: E402

for (W, H, mode) in ((160, 120, 0), (100, 75, 1)):
    depth = synth.depth_frames(2, W=W, H=H, seed=7)
    target = synth.target_frames(2, W=W, H=H, seed=8, blobs=6)
    sb = tod_b200.SceneBuilder(width=W, height=H, max_batch=2, weights_mode=mode)
    got = sb.append_batch(depth, target)
    s = sb.materialize(1)
    s.balls[:3] = [(W // 3, H // 3), (W // 2, H // 2), (5, 5)]
    p = tod_b200.modify_path(s, height_px=H)
    print("scene %dx%d mode %d: map sum %d, path %s" % (W, H, mode, int(got["map"].sum()), None if p is None else len(p.directions)))
if "--scene-only" not in sys.argv:
    _, small = synth_model.ensure_models()
    y = tod_b200.Yolact.init(small, max_tiles=2)
    r = y.infer_tiles(synth.rgb_tiles(2, S=64, seed=3), cell_classes=True)
    print("yolact small: outputs", [int(o.sum()) for o in r["outputs"]])

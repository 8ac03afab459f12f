#!/usr/bin/env python
"""Runs the scene path once on a small synthetic batch: the short program `ncu --set full` profiles.
    python tools/scene_probe.py [frames]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import tod_b200  # noqa: E402
from tests import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
depth = np.tile(synth.depth_frames(8, seed=3), (max(1, n // 8), 1, 1))[:n]
sb = tod_b200.SceneBuilder(max_batch=n)
for _ in range(2):
    sb.append_batch(depth, np.zeros_like(depth), want=("map",))
print("stamp %.3f ms, weights %.3f ms for %d frames" % (*sb.last_kernel_ms(), n))

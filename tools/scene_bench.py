#!/usr/bin/env python
"""Device-resident timing of the scene path (BASELINE config 3: 256 frames of 640x480): total ms per batch, and the stamp /
weights shares, for the tuning knobs read at handle creation (TOD_SCENE_CHUNK, TOD_WEIGHTS_VARIANT, TOD_STAMP_IMPL).
    python tools/scene_bench.py [frames] [chunk,chunk,...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import tod_b200  # noqa: E402
from tests import synth  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 256
chunks = [int(c) for c in sys.argv[2].split(",")] if len(sys.argv) > 2 else [int(os.environ.get("TOD_SCENE_CHUNK", "0"))]
depth = torch.from_numpy(np.tile(synth.depth_frames(8, seed=3), (nb // 8, 1, 1)).astype(np.int16)).cuda()
target = torch.zeros_like(depth)
npx = 640 * 480
o_map = torch.empty((nb, npx), dtype=torch.int32, device="cuda")
o_w = torch.empty((nb, npx, 4), dtype=torch.float32, device="cuda")
o_c0, o_c1 = torch.empty_like(o_w), torch.empty_like(o_w)
o_b = torch.empty((nb, 100, 4), dtype=torch.float32, device="cuda")
st = torch.cuda.Stream()
for chunk in chunks:
    os.environ["TOD_SCENE_CHUNK"] = str(chunk)
    sb = tod_b200.SceneBuilder(max_batch=nb)

    def step(stream):
        sb.append_batch_device(depth.data_ptr(), target.data_ptr(), nb, o_map.data_ptr(), o_w.data_ptr(), o_c0.data_ptr(), o_c1.data_ptr(), o_b.data_ptr(), stream)

    for _ in range(3):
        step(st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record()
        for _ in range(10):
            step(st.cuda_stream)
        e1.record()
    torch.cuda.synchronize()
    total = e0.elapsed_time(e1) / 10
    step(None)
    torch.cuda.synchronize()
    s_ms, w_ms = sb.last_kernel_ms()
    print("chunk %4d: total %.3f ms / %d frames (%.0f frames/s, %.2f TB/s algorithmic), stamp %.3f, weights %.3f" % (
        chunk, total, nb, nb / (total * 1e-3), nb * (56 * npx + 1600) / (total * 1e-3) / 1e12, s_ms, w_ms))
    del sb

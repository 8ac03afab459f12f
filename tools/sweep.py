#!/usr/bin/env python
"""BASELINE config 5 on one GPU: fused RGB-D pipeline (320x240 frames: classify -> u16 target -> point cloud + weights, all on
the device) over batch sizes 1 .. 4096 frames.  Batches above 512 frames are run as consecutive 512-frame calls on one handle
(1024 tiles of activations = 20 GB).  Prints one JSON line per batch size.

    python tools/sweep.py [--max 4096] > profiles/sweep_rNN.jsonl
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import tod_b200  # noqa: E402
from oracle import synth_model  # noqa: E402  (model file generator only)
from tests import synth  # noqa: E402

W, H, CHUNK = 320, 240, 512


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max", type=int, default=4096)
    a = ap.parse_args()
    full, _ = synth_model.ensure_models()
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    base_rgb = synth.rgb_frames(8, W=W, H=H, seed=5)
    base_depth = synth.depth_frames(8, W=W, H=H, seed=3)
    b = 1
    while b <= a.max:
        chunk = min(b, CHUNK)
        y = tod_b200.Yolact.init(full, max_tiles=2 * chunk)
        sb = tod_b200.SceneBuilder(width=W, height=H, max_batch=chunk)
        reps = (chunk + 7) // 8
        fr = torch.from_numpy(np.tile(base_rgb, (reps, 1))[:chunk].copy().view(np.int32)).cuda()
        dp = torch.from_numpy(np.tile(base_depth, (reps, 1, 1))[:chunk].copy().view(np.int16)).cuda()
        work = fr.clone()
        tg = torch.zeros((chunk, H, W), dtype=torch.int16, device="cuda")
        o_map = torch.empty((chunk, H * W), dtype=torch.int32, device="cuda")
        o_w = torch.empty((chunk, H * W, 4), dtype=torch.float32, device="cuda")
        o_c0, o_c1 = torch.empty_like(o_w), torch.empty_like(o_w)
        calls = (b + chunk - 1) // chunk

        def step():
            for _ in range(calls):
                work.copy_(fr)
                y.classify_device(work.data_ptr(), chunk, W, H, tg.data_ptr(), stream.cuda_stream)
                sb.append_batch_device(dp.data_ptr(), tg.data_ptr(), chunk, o_map.data_ptr(), o_w.data_ptr(), o_c0.data_ptr(), o_c1.data_ptr(), None,
                                       stream.cuda_stream)

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        iters = max(2, min(50, 2048 // b))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(json.dumps({"workload": "fused RGB-D 320x240 (configs[3]/[4])", "batch_frames": b, "calls_per_batch": calls, "ms_per_batch": ms,
                          "frames_per_sec": b / (ms * 1e-3)}), flush=True)
        del y, sb, fr, dp, work, tg, o_map, o_w, o_c0, o_c1
        torch.cuda.empty_cache()
        b *= 2


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Timeline of one YOLACT step on the executor's lane streams (tod_yolact_trace_steps): end time of every launch, the
lane it ran on, and - from the isolated per-launch durations of profile_ops - how long each launch waited.

    python tools/timeline.py [--tiles 64] > profiles/timeline_rNN.txt
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("TOD_PROFILE_REPS", "10")
import numpy as np  # noqa: E402

import tod_b200  # noqa: E402
from oracle import synth_model  # noqa: E402
from tests import synth  # noqa: E402

NAMES = {0: "ADD", 2: "CONCAT", 3: "CONV_2D", 4: "DEPTHWISE", 19: "RELU", 22: "RESHAPE", 23: "RESIZE", 28: "TANH", 34: "PAD", 114: "QUANTIZE"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiles", type=int, default=64)
    a = ap.parse_args()
    full, _ = synth_model.ensure_models()
    y = tod_b200.Yolact.init(full, max_tiles=a.tiles)
    y.infer_tiles(synth.rgb_tiles(a.tiles), outputs=False, tile_classes=False, detections=True, float_masks=False)
    iso, _ = y.profile_ops(a.tiles)
    end, lanes, kinds = y.trace_steps(a.tiles)
    ns = len(iso)
    print("# tiles=%d: end = end of the launch (us from step start, events on the lane streams, no graph); iso = isolated "
          "duration; start~ = end - iso; lane-gap = start~ minus the previous end on the same lane" % a.tiles)
    last_end = {}
    order = np.argsort(end[:ns], kind="stable")
    print("%4s %-14s %4s %9s %8s %9s %9s" % ("step", "op", "lane", "end", "iso", "start~", "lane-gap"))
    busy = 0.0
    for i in order:
        k = int(kinds[i])
        name = NAMES.get(k & 0xFFF, str(k)) + ("/tc" if k & 0x1000 else "")
        e, d = float(end[i]) * 1e3, float(iso[i]) * 1e3
        gap = e - d - last_end.get(int(lanes[i]), 0.0)
        last_end[int(lanes[i])] = e
        busy += d
        print("%4d %-14s %4d %9.1f %8.1f %9.1f %9.1f" % (i, name, lanes[i], e, d, e - d, gap))
    for name, t in zip(("seg post", "boxes/nms/topk", "masks (step end)"), end[ns:]):
        print("%-19s      %9.1f" % (name, float(t) * 1e3))
    print("# sum of isolated durations %.1f us; last graph step ends %.1f us; step ends %.1f us" % (busy, float(end[:ns].max()) * 1e3, float(end[-1]) * 1e3))
    y.close()


if __name__ == "__main__":
    main()

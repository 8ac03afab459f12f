#!/usr/bin/env python
"""Per-step device time of the planned YOLACT graph (CUDA events around every launch, n tiles resident).

    python tools/profile_ops.py [--tiles 64] [--conv-impl 0] > profiles/ops_rNN.txt
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import oracle  # noqa: E402  (only to name the ops / shapes of the model file)
import tod_b200  # noqa: E402
from oracle import synth_model  # noqa: E402
from tests import synth  # noqa: E402

NAMES = {0: "ADD", 2: "CONCAT", 3: "CONV_2D", 4: "DEPTHWISE", 19: "RELU", 22: "RESHAPE", 23: "RESIZE", 28: "TANH", 34: "PAD", 114: "QUANTIZE"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiles", type=int, default=64)
    ap.add_argument("--conv-impl", type=int, default=0)
    a = ap.parse_args()
    full, _ = synth_model.ensure_models()
    y = tod_b200.Yolact.init(full, max_tiles=a.tiles, conv_impl=a.conv_impl)
    y.infer_tiles(synth.rgb_tiles(a.tiles), outputs=False, tile_classes=False)
    ms, kinds = y.profile_ops(a.tiles)
    ms2, _ = y.profile_ops(a.tiles)
    ms = np.minimum(ms, ms2)
    print("# tiles=%d conv_impl=%d total=%.3f ms  (tcgen05 convs: %d)" % (a.tiles, a.conv_impl, ms.sum(), y.stats()["tc_conv_layers"]))
    agg = {}
    for t, k in zip(ms, kinds):
        name = NAMES.get(int(k) & 0xFFF, str(k)) + ("/tcgen05" if k & 0x1000 else "") + ("/copy" if k & 0x2000 else "")
        agg.setdefault(name, [0, 0.0])
        agg[name][0] += 1
        agg[name][1] += float(t)
    for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("# %-20s x%-3d %8.3f ms  %5.1f%%" % (name, c, t, 100 * t / ms.sum()))
    for i, (t, k) in enumerate(zip(ms, kinds)):
        print("%3d %-18s %8.4f ms" % (i, NAMES.get(int(k) & 0xFFF, str(k)) + ("/tc" if k & 0x1000 else ""), t))


if __name__ == "__main__":
    main()

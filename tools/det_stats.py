#!/usr/bin/env python
"""Candidate statistics of the detection head on the bench workload (how many priors per class pass the
confidence threshold, survivors, detections) - sizing input for the Fast-NMS kernels.  python tools/det_stats.py [tiles]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import tod_b200  # noqa: E402
from oracle import synth_model  # noqa: E402
from tests import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
full, _ = synth_model.ensure_models()
y = tod_b200.Yolact.init(full, max_tiles=n)
res = y.infer_tiles(synth.rgb_tiles(n, seed=2), detections=True, masks=False)
info = [y.output_info(k) for k in range(5)]
cls_i = [k for k in range(5) if info[k]["shape"][3] == 81 and info[k]["shape"][1] == 1][0]
q = res["outputs"][cls_i].reshape(n, -1, 81).astype(np.float32)
sc = info[cls_i]["scale"]
e = np.exp(sc * (q - q.max(axis=2, keepdims=True)))
p = e / e.sum(axis=2, keepdims=True)
cand = (p[:, :, 1:] > 0.05).sum(axis=1)  # [tile, class]
print("candidates per (tile, class): mean %.1f max %d; classes with >0: %.1f / 80; per tile total: %s" % (
    cand.mean(), cand.max(), (cand > 0).sum(axis=1).mean(), cand.sum(axis=1)[:8]))
print("detections per tile:", {k: (v["count"][:8] if isinstance(v, dict) and "count" in v else None) for k, v in res.items() if k.startswith("det")})
hist = np.histogram(cand.reshape(-1), bins=[0, 1, 16, 32, 64, 128, 200, 256, 512, 1024, 4096, 1 << 20])[0]
print("list-length histogram [0,1,16,32,64,128,200,256,512,1024,4096,inf):", hist.tolist())
print("sum n^2/2 over lists: %.0f per tile; sum min(n,200)^2/2: %.0f per tile" % ((cand.astype(np.float64) ** 2 / 2).sum() / n, (np.minimum(cand, 200).astype(np.float64) ** 2 / 2).sum() / n))

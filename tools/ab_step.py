#!/usr/bin/env python
"""Step time of the YOLACT graph (64 resident tiles, detections on) for A/B runs of two library builds on the same box:

    for lib in a.so b.so a.so b.so; do TOD_B200_LIB=$lib python tools/ab_step.py; done
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import tod_b200  # noqa: E402
from oracle import synth_model  # noqa: E402
from tests import synth  # noqa: E402

tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
full, _ = synth_model.ensure_models()
OPTS = {k[8:].lower(): int(v) for k, v in os.environ.items() if k.startswith("TOD_OPT_")}  # e.g. TOD_OPT_USE_PDL=0
y = tod_b200.Yolact.init(full, max_tiles=tiles, **OPTS)
d = torch.from_numpy(synth.rgb_tiles(tiles)).cuda()
ts = torch.cuda.Stream()  # events must sit on the stream the library launches on (a NULL stream means the handle's own)
torch.cuda.set_stream(ts)
stream = ts.cuda_stream
for _ in range(10):
    y.infer_tiles_device(d.data_ptr(), tiles, stream)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        y.infer_tiles_device(d.data_ptr(), tiles, stream)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / steps)
print("%s: %.4f ms/step (%d tiles, best of 3 x %d)" % (os.environ.get("TOD_B200_LIB", "default"), best, tiles, steps))
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 3
if depth > 1:  # batches in flight, as bench.py's headline: `depth` handles take alternate steps on their own streams
    y.close()
    popts = dict(OPTS)
    popts.setdefault("batches_in_flight", depth)
    ys = [tod_b200.Yolact.init(full, max_tiles=tiles, **popts) for _ in range(depth)]
    pst = [torch.cuda.Stream() for _ in range(depth)]
    for k in range(4 * depth):
        ys[k % depth].infer_tiles_device(d.data_ptr(), tiles, pst[k % depth].cuda_stream)
    torch.cuda.synchronize()
    pbest = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for ps in pst:
            ps.wait_event(e0)
        for k in range(steps):
            ys[k % depth].infer_tiles_device(d.data_ptr(), tiles, pst[k % depth].cuda_stream)
        for ps in pst:
            ts.wait_stream(ps)
        e1.record()
        torch.cuda.synchronize()
        pbest = min(pbest, e0.elapsed_time(e1) / steps)
    print("%s: %.4f ms/step with %d batches in flight" % (os.environ.get("TOD_B200_LIB", "default"), pbest, depth))
    for yy in ys:
        yy.close()
else:
    y.close()

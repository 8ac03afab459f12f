import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import tod_b200
from oracle import synth_model
from tests import synth
full, _ = synth_model.ensure_models()
def run(nh, per):
    ys = [tod_b200.Yolact.init(full, max_tiles=per) for _ in range(nh)]
    streams = [torch.cuda.Stream() for _ in range(nh)]
    tiles = [torch.from_numpy(synth.rgb_tiles(per, seed=2 + i)).cuda() for i in range(nh)]
    def step():
        for y, s, t in zip(ys, streams, tiles):
            y.infer_tiles_device(t.data_ptr(), per, s.cuda_stream)
    for _ in range(5): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    import time; t0 = time.perf_counter()
    for _ in range(20): step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 20
    print("handles=%d x %d tiles: %.3f ms per %d tiles -> %.0f frames/s" % (nh, per, dt * 1e3, nh * per, nh * per / 2 / dt))
run(1, 64); run(2, 32); run(4, 16); run(2, 64); run(1, 128); run(1, 32)

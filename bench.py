#!/usr/bin/env python
"""bench.py - headline benchmark of the B200-native perception hot path.

Workload (BASELINE.json configs[1]): the FRC int8 YOLACT graph + decode / Fast-NMS / mask assembly on a
synthetic RGB batch of 64 tiles (224x224x3) per GPU.  One *step* = one pass of that hot path over one
batch.  metric = YOLACT frames/s, a camera frame being the two 224x224 tiles the reference cuts from it
(src/yolact.rs:213-217), so 64 tiles = 32 frames.  Independent frames are sharded over ranks with no
collective (weak scaling: every rank runs its own 64 tiles).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  `value` is measured with inputs resident in HBM; `e2e` goes through the
reference-facing C-ABI call with host buffers (H2D + D2H inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TILES_PER_STEP = 64
TILES_PER_FRAME = 2
SCENE_BATCH = 256  # BASELINE.json configs[2]


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm),
                "power_w_max": max(pw) if pw else None}


def bench_config(n, world, depth, e2e_depth=None):
    """config of the line: the same dict for both arms (the reference arm times a bounded sample of this workload)"""
    return {"workload": "FRC_model int8 YOLACT full graph + decode/Fast-NMS/mask assembly, synthetic RGB batch %d tiles per GPU (configs[1])" % n,
            "tiles_per_step": n, "frames_per_step": n // TILES_PER_FRAME, "tile": "224x224x3 u8",
            "graph": "synthetic FRC_model.tflite stand-in with the reference's operator histogram (real blob missing), 5.62 GMAC/tile",
            "parallelism": "frame-sharded x%d, no collective" % world,
            "pipeline": ("%d batches in flight: %d handles take alternate %d-tile steps on their own streams (frame-loop double buffering); "
                         "the end-to-end leg keeps %d in flight (the extra ones hide the host copies); single_stream = one handle, steps back to back"
                         % (depth, depth, n, e2e_depth or depth)) if depth > 1 else "1 (one handle, steps back to back)",
            "l2": "activation working set ~%.0f MB per step (19.5 MB/tile) > 126 MB L2; no flush needed" % (19.5 * n)}


def _crc(a):
    import zlib
    import numpy as np
    return zlib.crc32(np.ascontiguousarray(a).tobytes())


def self_check(tod_b200, y, full, n, rank, world, timed_is_golden_input):
    """The bytes behind the number.  (1) rank 0's timed input is tests.synth.rgb_tiles(64, seed=2), the input the CPU oracle's
    committed CRCs (tests/golden/config2_oracle.json, tools/make_golden.py) were computed on: the outputs, class grids and
    detections the LAST TIMED STEP left on the device are fetched and compared tile by tile.  (2) every rank then runs that
    same input once and the digests are compared across ranks: a frame's bytes do not depend on the GPU it ran on
    (SURVEY 8e).  Never raises: a mismatch is reported in the line."""
    import ctypes as C
    import hashlib
    import numpy as np
    from tests import synth
    out = {"golden": "tests/golden/config2_oracle.json (CPU oracle, all %d tiles: 5 outputs + class grid + NMS keep indices / classes / scores / boxes)" % min(n, 64)}
    try:
        g = json.load(open(os.path.join(ROOT, "tests", "golden", "config2_oracle.json")))
        model_ok = hashlib.sha256(open(full, "rb").read()).hexdigest() == g["model_sha256"]
        out["model_matches_golden"] = model_ok

        def digest_of_last_call(m):
            outs = [y.fetch_output(k, m) for k in range(5)]
            det, keep = y._alloc_dets(m, True)
            det.masks = None
            det.masks_bin = None
            tod_b200._lib.check(tod_b200.lib().tod_yolact_fetch_detections(y._h, m, C.byref(det)))
            cells = y.fetch_tile_classes(m)[:, ::8, ::8]
            bad, crcs = 0, []
            for t in range(m):
                c = keep["count"][t]
                rec = {"out": [_crc(o[t]) for o in outs], "cells": _crc(cells[t]), "n_det": int(c), "prior": _crc(keep["priors"][t, :c]),
                       "cls": _crc(keep["classes"][t, :c]), "score": _crc(keep["scores"][t, :c]), "box": _crc(keep["boxes"][t, :c])}
                crcs.append(rec)
                if model_ok and t < len(g["tiles"]):
                    w = g["tiles"][t]
                    bad += int(any(rec[k] != w[k] for k in ("out", "cells", "n_det", "prior", "cls", "score", "box")))
            return bad, _crc(np.array([[r["cells"], r["prior"], r["score"], r["box"]] + r["out"] for r in crcs], np.uint32))

        m = min(n, 64)
        if rank == 0 and timed_is_golden_input:
            bad, dig = digest_of_last_call(m)
            out["timed_step_tiles_checked"] = m
            out["timed_step_tiles_differing_from_oracle"] = bad if model_ok else None
        tiles = synth.rgb_tiles(64, seed=2)[:m]
        y.infer_tiles(tiles, outputs=False, tile_classes=False, detections=True, float_masks=False)
        bad2, dig2 = digest_of_last_call(m)
        out["rerun_tiles_differing_from_oracle"] = bad2 if model_ok else None
        out["digest"] = int(dig2)
        if world > 1:
            from tests import dist_helpers
            out["ranks_agree"] = bool(dist_helpers.all_equal_over_ranks(dig2))
        out["ok"] = bool((not model_ok or (bad2 == 0 and out.get("timed_step_tiles_differing_from_oracle", 0) in (0, None))) and out.get("ranks_agree", True))
    except Exception as e:  # the check must never cost the measurement
        out["error"] = str(e)[:300]
        out["ok"] = False
    return out


def run_reference(args, rank):
    """The reference's own CPU implementation of the path, restated (the Rust + TFLite original cannot be
    built here): oracle int8 graph + literal post-processing + detection on the host cores."""
    if rank != 0:
        return
    import numpy as np
    import oracle
    from oracle import synth_model
    from tests import synth
    full, _ = synth_model.ensure_models()
    cores = os.cpu_count() or 1
    m = oracle.Model(full)
    outs = [m.tensor_info(t) for t in m.outputs]
    sample_tiles = 2  # one camera frame per step: a bounded sample of the 64-tile batch
    tiles = synth.rgb_tiles(sample_tiles, seed=2)

    def step():
        for t in range(sample_tiles):
            m.invoke(tiles[t], threads=cores)
            o = [m.tensor(ti) for ti in m.outputs]
            oracle.postprocess_tile(o[4], outs[4]["scale"], outs[4]["zero_point"], 0)
            oracle.detect(o[1], (outs[1]["scale"], outs[1]["zero_point"]), o[0], (outs[0]["scale"], outs[0]["zero_point"]),
                          o[2], (outs[2]["scale"], outs[2]["zero_point"]), o[3], (outs[3]["scale"], outs[3]["zero_point"]))

    for _ in range(min(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    fps = args.steps * sample_tiles / TILES_PER_FRAME / dt
    line = {
        "impl": "reference", "metric": "yolact_frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int8", "data": "synthetic",
        "config": bench_config(args.tiles, args.gpus, max(1, args.pipeline), max(args.pipeline, args.e2e_pipeline) if args.pipeline > 1 else 1),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": "%d tiles (1 camera frame) of the %d-tile step per timed step, through oracle/ (int8 graph with TFLite reference-kernel loop nests, literal postprocess, decode / Fast-NMS / masks; OpenMP over %d threads)" % (sample_tiles, args.tiles, cores)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--tiles", type=int, default=TILES_PER_STEP)
    ap.add_argument("--no-scene", action="store_true", help="skip the point-cloud side measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--conv-impl", type=int, default=0)
    ap.add_argument("--pipeline", type=int, default=3,
                    help="batches in flight of the device-resident leg: handles that take alternate steps on their own streams (1 = one handle, steps back to back)")
    ap.add_argument("--e2e-pipeline", type=int, default=5,
                    help="batches in flight of the end-to-end leg (host buffers): two more than the device-resident leg hide the copies")
    ap.add_argument("--fused", type=int, default=512, help="frames of the fused 320x240 RGB-D side measurement (0 = skip)")
    ap.add_argument("--sustain", type=float, default=2.0, help="seconds of the sustained side figure (0 = skip)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import tod_b200
    from oracle import synth_model  # model *file* generator only (the real blob is absent from the reference)
    from tests import synth

    if not torch.cuda.is_available() or tod_b200.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    full, _ = synth_model.ensure_models()
    n = args.tiles
    y = tod_b200.Yolact.init(full, device=local_rank, max_tiles=n, conv_impl=args.conv_impl)
    st = y.stats()
    # a real (non-default) stream: the C ABI treats a NULL stream as "the handle's own stream", and CUDA events
    # recorded by torch must sit on the stream the kernels are launched on
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    tiles_h = torch.from_numpy(synth.rgb_tiles(n, seed=2 + rank)).pin_memory()
    tiles_d = tiles_h.cuda()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def step():
        y.infer_tiles_device(tiles_d.data_ptr(), n, stream)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    barrier()
    frames_per_step = n // TILES_PER_FRAME
    value = world * frames_per_step * args.steps / (ms * 1e-3)
    check = self_check(tod_b200, y, full, n, rank, world, timed_is_golden_input=(n == 64))
    barrier()

    # ---- end to end through the reference-facing C-ABI call with host buffers
    out_bufs = y._alloc_dets(n, True)
    GH, GW = y.outputs[4]["shape"][1], y.outputs[4]["shape"][2]
    # the literal postprocess result as the 28x28 (class, id) grid: tile_classes is its 8x8 replication (yolact.rs:127-128),
    # 64x the bytes for the same information
    tc_h = torch.empty((n, GH, GW), dtype=torch.int32).pin_memory()
    det, keep = out_bufs
    keep_pinned = {}
    import ctypes as C
    for k in ("count", "boxes", "scores", "classes", "priors", "masks_bits"):
        tt = torch.from_numpy(keep[k].view(np.int32) if keep[k].dtype == np.uint32 else keep[k]).pin_memory()
        keep_pinned[k] = tt
        setattr(det, k, tt.data_ptr())
    det.masks = None
    det.masks_bin = None  # binary masks come back bit-packed (masks_bits): an eighth of the bytes
    d2h = sum(t.numel() * t.element_size() for t in keep_pinned.values()) + tc_h.numel() * 4
    h2d = tiles_h.numel()
    lib = tod_b200.lib()

    def e2e_step():
        tod_b200._lib.check(lib.tod_yolact_infer_tiles_cells(y._h, tiles_h.data_ptr(), n, None, None, tc_h.data_ptr(), C.byref(det)))

    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = world * frames_per_step * args.steps / e2e_s
    single = {"value": value, "ms_per_step": ms / args.steps, "e2e": e2e}

    # ---- two (or more) batches in flight: the frame loop's double buffering.  Every step is still one full batch through one
    # handle; consecutive steps go to alternate handles on their own streams, so the latency-bound tail of one batch (small
    # pyramid levels, Fast-NMS, mask assembly) overlaps the backbone of the next, and in the end-to-end leg one batch's copies
    # overlap the other's kernels (one host thread per handle: the C-ABI call blocks, ctypes drops the GIL).
    depth = max(1, args.pipeline)
    if depth > 1:
        import threading
        # handles that share the GPU say so (tod_yolact_options::batches_in_flight): their convolution launches give up SMs to each other
        ys = [tod_b200.Yolact.init(full, device=local_rank, max_tiles=n, conv_impl=args.conv_impl, batches_in_flight=depth) for _ in range(depth)]
        pstreams = [torch.cuda.Stream() for _ in range(depth)]
        ptiles = [tiles_d] + [torch.from_numpy(synth.rgb_tiles(n, seed=100 + 7 * k + rank)).cuda() for k in range(1, depth)]  # ys[0] runs the golden input

        def pstep(k):
            h = k % depth
            ys[h].infer_tiles_device(ptiles[h].data_ptr(), n, pstreams[h].cuda_stream)

        for k in range(2 * depth + max(args.warmup, 3)):
            pstep(k)
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()  # on tstream; every pipeline stream starts behind it and tstream ends behind all of them
        for ps in pstreams:
            ps.wait_event(p0)
        for k in range(args.steps):
            pstep(k)
        for ps in pstreams:
            tstream.wait_stream(ps)
        p1.record()
        torch.cuda.synchronize()
        pms = p0.elapsed_time(p1)
        if dist is not None:
            t = torch.tensor([pms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            pms = float(t.item())
        barrier()
        ms = pms
        value = world * frames_per_step * args.steps / (pms * 1e-3)
        # the bytes of the headline region: ys[0] ran the golden input in every one of its timed steps
        check_single = check
        check = self_check(tod_b200, ys[0], full, n, rank, world, timed_is_golden_input=(n == 64))
        check["single_stream_ok"] = bool(check_single.get("ok"))
        barrier()

        # sustained side figure: the same pipelined steps for >= 2 s under its own clock / power record (the headline region is
        # only steps x ~1 ms long)
        sustained = None
        if args.sustain > 0:
            sus_steps = max(args.steps, int(args.sustain * 1e3 / max(pms / args.steps, 1e-3)))
            sclk = ClockSampler(local_rank)
            if rank == 0:
                sclk.start()
            p0.record()
            for ps in pstreams:
                ps.wait_event(p0)
            for k in range(sus_steps):
                pstep(k)
            for ps in pstreams:
                tstream.wait_stream(ps)
            p1.record()
            torch.cuda.synchronize()
            sus_ms = p0.elapsed_time(p1)
            if dist is not None:
                t = torch.tensor([sus_ms], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                sus_ms = float(t.item())
            sustained = {"value": world * frames_per_step * sus_steps / (sus_ms * 1e-3), "unit": "frames/s", "steps": sus_steps, "seconds": sus_ms * 1e-3,
                         "clocks": sclk.stop() if rank == 0 else None}
            barrier()

        # end to end: one host thread per handle, each with its own pinned buffers
        e2e_depth = max(depth, args.e2e_pipeline)
        for k in range(depth, e2e_depth):
            ys.append(tod_b200.Yolact.init(full, device=local_rank, max_tiles=n, conv_impl=args.conv_impl, batches_in_flight=e2e_depth))
            ptiles.append(torch.from_numpy(synth.rgb_tiles(n, seed=100 + 7 * k + rank)).cuda())
        ctx = []
        for h in range(e2e_depth):
            d_h, keep_h = ys[h]._alloc_dets(n, True)
            pin = {}
            for kk in ("count", "boxes", "scores", "classes", "priors", "masks_bits"):
                tt = torch.from_numpy(keep_h[kk].view(np.int32) if keep_h[kk].dtype == np.uint32 else keep_h[kk]).pin_memory()
                pin[kk] = tt
                setattr(d_h, kk, tt.data_ptr())
            d_h.masks = None
            d_h.masks_bin = None
            ctx.append((ys[h], d_h, pin, torch.empty((n, GH, GW), dtype=torch.int32).pin_memory(), (tiles_h if h == 0 else ptiles[h].cpu().pin_memory())))
        counts = [args.steps // e2e_depth + (1 if h < args.steps % e2e_depth else 0) for h in range(e2e_depth)]

        def worker(h, reps):
            torch.cuda.set_device(local_rank)
            yy, d_h, _, tc_p, th = ctx[h]
            for _ in range(reps):
                tod_b200._lib.check(lib.tod_yolact_infer_tiles_cells(yy._h, th.data_ptr(), n, None, None, tc_p.data_ptr(), C.byref(d_h)))

        def run_threads(cs):
            ths = [threading.Thread(target=worker, args=(h, cs[h])) for h in range(e2e_depth)]
            for th in ths:
                th.start()
            for th in ths:
                th.join()

        run_threads([2] * e2e_depth)
        barrier()
        t0 = time.perf_counter()
        run_threads(counts)
        torch.cuda.synchronize()
        pe2e_s = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([pe2e_s], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            pe2e_s = float(t.item())
        e2e = world * frames_per_step * args.steps / pe2e_s
    clk = clocks.stop() if rank == 0 else None  # sampled over the timed regions (device-resident and end-to-end)

    # ---- roofline of the dominant kernel: conv_tc_fast_kernel (tcgen05 int8 implicit GEMM), every launch of one step.
    # achieved = algorithmic ops of those launches (2 * MACs, SURVEY 8d: 11.24 Gop/tile over the whole graph) / the sum of
    # their CUDA-event durations on the launching stream; peak = tcgen05 kind::i8 tensor-pipe peak measured in this
    # process (MEASURED_PEAKS.json has no int8 entry; 2 x its bf16 burst figure is reported beside it).
    peaks = _peaks()
    ms_ops, kinds = y.profile_ops(n)
    ms_ops2, _ = y.profile_ops(n)
    ms_ops = np.minimum(ms_ops, ms_ops2)
    macs_ops = y.step_macs()
    is_tc = (kinds & 0x1000) != 0
    tc_ms = float(ms_ops[is_tc].sum())
    tc_macs = float(macs_ops[is_tc].sum()) * n
    total_ops_ms = float(ms_ops.sum())
    i8_peak = tod_b200._lib.i8_mma_peak(n_mma=20000, iters=5, device=local_rank)
    achieved = 2.0 * tc_macs / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "conv_tc_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
    big = int(np.argmax(np.where(is_tc, macs_ops, 0)))
    # The timed region runs `depth` batches concurrently, and the convolution launches are sized for that (balanced grids, at
    # least two tiles per CTA: a launch alone on the GPU takes longer, two of them side by side finish sooner).  The duration a
    # launch costs IN the timed region's execution mode is measured the same way as the region itself: the same handles'
    # graphs with everything but the tcgen05 convolutions left out (TOD_DIAG_SKIP = 27 at handle creation: no depthwise /
    # resize / stem / detection tail), `depth` batches in flight, CUDA events on the launching streams.  The serial figure
    # (every launch alone, CUDA events, summed - round 1's definition) stays beside it as `isolated`.
    conc_ms = None
    try:
        os.environ["TOD_DIAG_SKIP"] = "27"
        cys = [tod_b200.Yolact.init(full, device=local_rank, max_tiles=n, conv_impl=args.conv_impl, batches_in_flight=depth) for _ in range(depth)]
    finally:
        os.environ.pop("TOD_DIAG_SKIP", None)
    cstreams = [torch.cuda.Stream() for _ in range(depth)]
    for k in range(4 * depth):
        cys[k % depth].infer_tiles_device(tiles_d.data_ptr(), n, cstreams[k % depth].cuda_stream)
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    csteps = max(args.steps, 50)
    c0.record()
    for cs_ in cstreams:
        cs_.wait_event(c0)
    for k in range(csteps):
        cys[k % depth].infer_tiles_device(tiles_d.data_ptr(), n, cstreams[k % depth].cuda_stream)
    for cs_ in cstreams:
        tstream.wait_stream(cs_)
    c1.record()
    torch.cuda.synchronize()
    conc_ms = c0.elapsed_time(c1) / csteps
    for cy in cys:
        cy.close()
    conc_achieved = 2.0 * tc_macs / (conc_ms * 1e-3) / 1e12
    n_tc = max(1, int(is_tc.sum()))
    roofline = {"bound": "tensor", "achieved": conc_achieved, "peak": i8_peak, "unit": "TOP/s", "frac": conc_achieved / i8_peak if i8_peak else None,
                "traffic": traffic, "traffic_source": traffic_src,
                "kernel": "conv_tc_fast_kernel / conv_tc_pair_kernel / conv_tc_flc_kernel: all %d tcgen05 conv launches of one step" % n_tc,
                "method": "achieved = algorithmic ops of the %d launches / their time in the timed region's execution mode: the handles' graphs with only the "
                          "tcgen05 convolutions (+ the 64-CTA literal segmentation pass) left in, %d batches in flight, CUDA events on the launching streams "
                          "(%.4f ms per step); `isolated` = every launch alone, CUDA events, summed (round 1's definition); `step_level` = the same ops / "
                          "the full step" % (n_tc, depth, conc_ms),
                "launches": n_tc, "avg_launch_ms": conc_ms / n_tc, "conv_only_ms_per_step": conc_ms,
                "isolated": {"achieved": achieved, "frac": achieved / i8_peak if i8_peak else None, "tc_ms_per_step": tc_ms, "avg_launch_ms": tc_ms / n_tc},
                "step_level": {"achieved": 2.0 * tc_macs / (ms / args.steps * 1e-3) / 1e12, "frac": 2.0 * tc_macs / (ms / args.steps * 1e-3) / 1e12 / i8_peak if i8_peak else None},
                "algorithmic_gop_per_launch": 2.0 * tc_macs / max(1, int(is_tc.sum())) / 1e9,
                "peak_source": "tcgen05.mma.kind::i8 issue-only micro-benchmark (tod_i8_mma_peak) measured in this run; MEASURED_PEAKS.json has no int8 entry (2 x its bf16 burst = %.0f TOP/s, %s)" % (2.0 * peaks["bf16"], peaks["src"]),
                "largest_launch": {"gop": 2.0 * float(macs_ops[big]) * n / 1e9, "ms": float(ms_ops[big]),
                                   "top_s": 2.0 * float(macs_ops[big]) * n / (float(ms_ops[big]) * 1e-3) / 1e12 if ms_ops[big] > 0 else None},
                "tc_ms_per_step": tc_ms, "all_ops_ms_per_step": total_ops_ms, "tc_conv_layers": st["tc_conv_layers"],
                "hbm": {"algorithmic_bytes_per_step": 19.5e6 * n, "ms_per_step": ms / args.steps,
                        "achieved_GBps": 19.5e6 * n / (ms / args.steps * 1e-3) / 1e9, "peak_GBps": peaks["hbm"]}}

    line = {
        "metric": "yolact_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8", "data": "synthetic",
        "config": bench_config(n, world, depth, max(depth, args.e2e_pipeline) if depth > 1 else 1),
        "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "call": "tod_yolact_infer_tiles_cells (host tiles in; 28x28 class grids + detections + bit-packed binary masks out)"},
        "single_stream": {"value": single["value"], "ms_per_step": single["ms_per_step"], "e2e_value": single["e2e"], "unit": "frames/s"},
        "gpu_launches": int(st["launches_per_call"] * args.steps),
        "check": check,
        "sustained": sustained if depth > 1 else None,
        "roofline": roofline,
        "clocks": clk,
        "tiles_per_sec": value * TILES_PER_FRAME,
    }

    # ---- detection tail (decode / Fast-NMS / top-k / mask assembly): HBM-side figure of BASELINE's metric.  Timed on the lane
    # streams without the graph (tod_yolact_trace_steps): from the end of the last graph layer to the end of mask assembly.
    try:
        t_end, _, _ = y.trace_steps(n)
        ns_steps = len(ms_ops)
        tail_ms = float(t_end[-1] - t_end[:ns_steps].max())
        det_bytes = n * (3147 * (81 + 4 + 32) + 56 * 56 * 32) + n * 100 * 56 * 56 * (4 + 1) + n * 100 * (56 * 56 // 8)  # heads + prototypes in; float + byte + bit masks out
        line["detection_tail"] = {"ms": tail_ms, "algorithmic_bytes": int(det_bytes), "GBps": det_bytes / (tail_ms * 1e-3) / 1e9,
                                  "frac_of_hbm": det_bytes / (tail_ms * 1e-3) / 1e9 / peaks["hbm"],
                                  "note": "decode + Fast-NMS + top-k + mask assembly after the last graph layer, stream-ordered (no graph); instruction-bound, see DESIGN 5.3"}
    except Exception as e:
        line["detection_tail"] = {"error": str(e)[:200]}

    def rank_max(v):
        if dist is None:
            return float(v)
        t = torch.tensor([float(v)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def host_timed(fn, reps, warm=2):
        """wall-clock seconds of `reps` blocking C-ABI calls (host buffers in and out), max over ranks"""
        for _ in range(warm):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return rank_max(time.perf_counter() - t0)

    # ---- scene path (configs[2]): depth -> point cloud + weights; every rank runs its own 256 frames (weak scaling)
    if not args.no_scene:
        try:
            nb = SCENE_BATCH
            sb = tod_b200.SceneBuilder(device=local_rank, max_batch=nb)
            base = synth.depth_frames(8, seed=3 + rank)
            depth = torch.from_numpy(np.tile(base, (nb // 8, 1, 1)).astype(np.int16)).cuda()
            target = torch.zeros_like(depth)
            npx = 640 * 480
            o_map = torch.empty((nb, npx), dtype=torch.int32, device="cuda")
            o_w = torch.empty((nb, npx, 4), dtype=torch.float32, device="cuda")
            o_c0, o_c1 = torch.empty_like(o_w), torch.empty_like(o_w)
            o_b = torch.empty((nb, 100, 4), dtype=torch.float32, device="cuda")

            def sstep():
                sb.append_batch_device(depth.data_ptr(), target.data_ptr(), nb, o_map.data_ptr(), o_w.data_ptr(), o_c0.data_ptr(), o_c1.data_ptr(), o_b.data_ptr(), stream)

            for _ in range(3):
                sstep()
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            s0.record()
            for _ in range(reps):
                sstep()
            s1.record()
            torch.cuda.synchronize()
            sms = rank_max(s0.elapsed_time(s1) / reps)
            bytes_per_frame = 56 * npx + 1600
            sb.append_batch_device(depth.data_ptr(), target.data_ptr(), nb, o_map.data_ptr(), o_w.data_ptr(), o_c0.data_ptr(), o_c1.data_ptr(), o_b.data_ptr(), None)
            stamp_ms, weights_ms = sb.last_kernel_ms()
            scene = {"workload": "640x480 depth -> point cloud + weights, batch %d per GPU (configs[2]), device-resident" % nb,
                     "frames_per_sec": world * nb / (sms * 1e-3), "ms_per_batch": sms,
                     "algorithmic_GBps_per_gpu": nb * bytes_per_frame / (sms * 1e-3) / 1e9, "hbm_peak_GBps": peaks["hbm"],
                     "frac_of_hbm": nb * bytes_per_frame / (sms * 1e-3) / 1e9 / peaks["hbm"], "stamp_ms": stamp_ms, "weights_ms": weights_ms,
                     "weights_GBps": nb * 52 * npx / (weights_ms * 1e-3) / 1e9, "weights_frac_of_hbm": nb * 52 * npx / (weights_ms * 1e-3) / 1e9 / peaks["hbm"],
                     "stamp_kernel": "stamp_pruned_kernel (one entry per column and landing row) + tile merge fused into weights_kernel",
                     "shader_stamps_per_sec": nb * npx * 400 / (stamp_ms * 1e-3)}
            del o_w, o_c0, o_c1, o_map
            # end to end through the drop-in call: host depth + target in, the reference's four read-backs out (scene.rs:246,257-259),
            # then the Scene conversion of one frame (scene.rs:312-327)
            eb = 32
            sbe = tod_b200.SceneBuilder(device=local_rank, max_batch=eb)
            h_depth = torch.from_numpy(np.tile(base, (eb // 8, 1, 1)).astype(np.int16)).pin_memory()
            h_target = torch.zeros_like(h_depth).pin_memory()
            h_map = torch.empty((eb, npx), dtype=torch.int32).pin_memory()
            h_w = torch.empty((eb, npx, 4), dtype=torch.float32).pin_memory()
            h_c0, h_c1 = torch.empty_like(h_w).pin_memory(), torch.empty_like(h_w).pin_memory()
            h_b = torch.empty((eb, 100, 4), dtype=torch.float32).pin_memory()
            sc_h = np.zeros(npx, np.float32), np.zeros((npx, 3), np.float32), np.zeros((100, 2), np.int32), np.zeros((npx, 8), np.float32)

            def scene_e2e():
                tod_b200._lib.check(lib.tod_scene_append_batch(sbe._h, h_depth.data_ptr(), h_target.data_ptr(), eb, h_map.data_ptr(), h_w.data_ptr(),
                                                               h_c0.data_ptr(), h_c1.data_ptr(), h_b.data_ptr()))
                tod_b200._lib.check(lib.tod_scene_materialize(sbe._h, eb - 1, sc_h[0].ctypes.data, sc_h[1].ctypes.data, sc_h[2].ctypes.data, sc_h[3].ctypes.data))

            es = host_timed(scene_e2e, 5)
            scene["e2e"] = {"value": world * eb * 5 / es, "unit": "frames/s", "h2d_bytes_per_step": int(eb * npx * 4),
                            "d2h_bytes_per_step": int(eb * (npx * 52 + 1600) + npx * 48 + 800),
                            "call": "tod_scene_append_batch (32 frames, host depth + target in, map + world + conn0 + conn1 + balls out) + tod_scene_materialize of one frame; PCIe-bound: 16 MB of read-back per frame as in the reference"}
            line["scene"] = scene
            del sb, sbe
        except Exception as e:
            line["scene"] = {"error": str(e)[:200]}

    # ---- fused RGB-D pipeline (configs[3] / [4]): 320x240 frames, classify -> u16 target -> point cloud, all on the device; every rank
    if args.fused:
        try:
            fb = args.fused
            W4, H4 = 320, 240
            y4 = tod_b200.Yolact.init(full, device=local_rank, max_tiles=2 * fb)
            sb4 = tod_b200.SceneBuilder(device=local_rank, width=W4, height=H4, max_batch=fb)
            fr = torch.from_numpy(np.tile(synth.rgb_frames(8, W=W4, H=H4, seed=5 + rank), (fb // 8, 1)).view(np.int32)).cuda()
            dp = torch.from_numpy(np.tile(synth.depth_frames(8, W=W4, H=H4, seed=3 + rank), (fb // 8, 1, 1)).view(np.int16)).cuda()
            tg = torch.zeros((fb, H4, W4), dtype=torch.int16, device="cuda")
            o_map = torch.empty((fb, H4 * W4), dtype=torch.int32, device="cuda")
            o_w = torch.empty((fb, H4 * W4, 4), dtype=torch.float32, device="cuda")
            o_c0, o_c1 = torch.empty_like(o_w), torch.empty_like(o_w)
            work = fr.clone()

            def fstep():
                work.copy_(fr)  # classify mutates the frames in place (yolact.rs:233)
                y4.classify_device(work.data_ptr(), fb, W4, H4, tg.data_ptr(), stream)
                sb4.append_batch_device(dp.data_ptr(), tg.data_ptr(), fb, o_map.data_ptr(), o_w.data_ptr(), o_c0.data_ptr(), o_c1.data_ptr(), None, stream)

            for _ in range(2):
                fstep()
            barrier()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(3):
                fstep()
            f1.record()
            torch.cuda.synchronize()
            fms = rank_max(f0.elapsed_time(f1) / 3)
            line["fused_rgbd"] = {"workload": "320x240 RGB-D frames: classify (2 tiles/frame) -> target -> point cloud + weights, batch %d per GPU, device-resident (configs[3]); configs[4] = this at N GPUs" % fb,
                                  "frames_per_sec": world * fb / (fms * 1e-3), "ms_per_batch": fms, "n_gpus": world}
            del y4, sb4, o_w, o_c0, o_c1, o_map, work
            # end to end behind the C ABI: tod_pool_rgbd_batch on this rank's GPU (3 handles, chunks of 32 frames), host frames + depth in,
            # classified frames + map + world + conn0 + conn1 out
            pool_depth = 4
            pool = tod_b200.Pool(full, devices=[local_rank], depth=pool_depth, max_tiles=64)
            pn = 192
            sp = tod_b200.default_params(width=W4, height=H4)
            pf = torch.from_numpy(np.tile(synth.rgb_frames(8, W=W4, H=H4, seed=5 + rank), (pn // 8, 1)).view(np.int32)).pin_memory()
            pd = torch.from_numpy(np.tile(synth.depth_frames(8, W=W4, H=H4, seed=3 + rank), (pn // 8, 1, 1)).view(np.int16)).pin_memory()
            pm = torch.empty((pn, H4 * W4), dtype=torch.int32).pin_memory()
            pw = torch.empty((pn, H4 * W4, 4), dtype=torch.float32).pin_memory()
            pc0, pc1 = torch.empty_like(pw).pin_memory(), torch.empty_like(pw).pin_memory()

            def rgbd_e2e():   # frames are classified in place; re-running on the previous result costs the same work
                tod_b200._lib.check(lib.tod_pool_rgbd_batch(pool._h, C.byref(sp), pf.data_ptr(), pd.data_ptr(), pn, pm.data_ptr(), pw.data_ptr(), pc0.data_ptr(), pc1.data_ptr(), None))

            es = host_timed(rgbd_e2e, 3, warm=1)
            line["fused_rgbd"]["e2e"] = {"value": world * pn * 3 / es, "unit": "frames/s", "h2d_bytes_per_step": int(pn * W4 * H4 * 6),
                                         "d2h_bytes_per_step": int(pn * W4 * H4 * (4 + 52)),
                                         "call": "tod_pool_rgbd_batch: 192 host frames + depth per call, %d handles behind the C ABI; 4.3 MB of read-back per frame (PCIe-bound)" % pool_depth}
            # the reference's own call: Yolact::classify on 640x480 frames, in place, host buffer (yolact.rs:39)
            cn = 32 * pool_depth   # one 32-frame chunk per handle and call: the warm-up call touches every handle's buffers
            cf = torch.from_numpy(np.tile(synth.rgb_frames(8, seed=5 + rank), (cn // 8, 1)).view(np.int32)).pin_memory()

            def classify_e2e():
                tod_b200._lib.check(lib.tod_pool_classify_batch(pool._h, cf.data_ptr(), cn, 640, 480))

            es = host_timed(classify_e2e, 3, warm=2)
            line["classify_e2e"] = {"value": world * cn * 3 / es, "unit": "frames/s", "h2d_bytes_per_step": int(cn * 640 * 480 * 4), "d2h_bytes_per_step": int(cn * 640 * 480 * 4),
                                    "call": "tod_pool_classify_batch: %d frames of 640x480 u32 per call, classified in place (Yolact::classify for every frame), %d handles behind the C ABI" % (cn, pool_depth)}
            # one synchronous caller, tiles: the pooled form of tod_yolact_infer_tiles_cells
            tn = 8 * n
            pt = torch.from_numpy(np.tile(synth.rgb_tiles(n, seed=2 + rank), (8, 1, 1, 1))).pin_memory()
            pdet, pkeep = y._alloc_dets(tn, True)
            ppin = {}
            for kk in ("count", "boxes", "scores", "classes", "priors", "masks_bits"):
                tt = torch.from_numpy(pkeep[kk].view(np.int32) if pkeep[kk].dtype == np.uint32 else pkeep[kk]).pin_memory()
                ppin[kk] = tt
                setattr(pdet, kk, tt.data_ptr())
            pdet.masks = None
            pdet.masks_bin = None
            pcells = torch.empty((tn, GH, GW), dtype=torch.int32).pin_memory()

            def pool_tiles():
                tod_b200._lib.check(lib.tod_pool_infer_tiles(pool._h, pt.data_ptr(), tn, None, None, pcells.data_ptr(), C.byref(pdet)))

            reps = max(3, args.steps // 8)
            es = host_timed(pool_tiles, reps, warm=2)
            line["e2e"]["one_caller"] = {"value": world * (tn // TILES_PER_FRAME) * reps / es, "unit": "frames/s",
                                         "call": "tod_pool_infer_tiles: %d tiles per blocking call, the library's own %d handles / host threads (no Python threads)" % (tn, pool_depth)}
            del pool
        except Exception as e:  # never let the side measurement break the headline line
            line.setdefault("fused_rgbd", {})["error"] = str(e)[:200]

    # ---- the Scene's consumer: path::modify_path on one materialised 640x480 scene, rank 0 only
    if not args.no_scene and rank == 0:
        try:
            sp1 = tod_b200.SceneBuilder(device=local_rank, max_batch=1, weights_mode=1)
            d1 = synth.depth_frames(1, seed=18)
            sp1.append_batch(d1, synth.target_frames(1, seed=19), want=())
            sc1 = sp1.materialize(0)
            sc1.balls[:3] = [(100, 60), (500, 200), (320, 400)]
            tod_b200.modify_path(sc1, device=local_rank)
            t0 = time.perf_counter()
            pth = tod_b200.modify_path(sc1, device=local_rank)
            line["path"] = {"ms_per_scene": 1e3 * (time.perf_counter() - t0), "directions": 0 if pth is None else int(len(pth.directions)),
                            "call": "tod_path_modify (host Scene in: 13.5 MB; GPU relaxation to the fixed point; (magnitude, rotation) list out), intent mode - the reference function panics on every input"}
            del sp1
        except Exception as e:
            line["path"] = {"error": str(e)[:200]}

    # ---- CPU baseline (oracle port of the reference's CPU path), rank 0 at N=1 only
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle
        cores = os.cpu_count() or 1
        m = oracle.Model(full)
        tl = synth.rgb_tiles(4, seed=2)
        m.invoke(tl[0], threads=cores)
        t0 = time.perf_counter()
        cnt = 0
        while time.perf_counter() - t0 < 10.0 and cnt < 64:
            m.invoke(tl[cnt % 4], threads=cores)
            cnt += 1
        cpu_fps = cnt / TILES_PER_FRAME / (time.perf_counter() - t0)
        line["cpu_baseline"] = {"value": cpu_fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": "%d tiles of the same batch through oracle/ int8 graph (reference-kernel loop nests, OpenMP %d threads), ~10 s" % (cnt, cores)}
        # also 1 thread (per-core figure) and 4 threads (the reference's set_num_threads(4), yolact.rs:34)
        for th in (4, 1):
            t0 = time.perf_counter()
            m.invoke(tl[0], threads=th)
            line["cpu_baseline"]["frames_per_sec_%d_threads" % th] = 1.0 / TILES_PER_FRAME / (time.perf_counter() - t0)
        # the point-cloud half of the path (pt_cloud.comp + pt_cloud_weights.comp restated in oracle/, one thread): SURVEY 8d
        if "scene" in line and "error" not in line["scene"]:
            dfr = synth.depth_frames(3, seed=3)
            tfr = np.zeros_like(dfr)
            op = oracle.scene_params()
            t0 = time.perf_counter()
            for f in range(3):
                mp, _ = oracle.pt_cloud(dfr[f], tfr[f], op)
                oracle.pt_cloud_weights(mp, op)
            line["scene"]["cpu_baseline_frames_per_sec"] = 3.0 / (time.perf_counter() - t0)
            line["scene"]["cpu_baseline_note"] = "oracle port of both shaders, 640x480, 1 host thread, 3 frames"
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

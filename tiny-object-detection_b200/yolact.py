"""Host-side mirror of the reference's `Yolact` (src/yolact.rs:13-41) over the C ABI.

    Yolact.init(path)            <- Yolact::init()                  (yolact.rs:17-37)
    Yolact.classify(frame)       <- Yolact::classify(&mut [u32])    (yolact.rs:39-41), in place
    Yolact.infer_tiles(tiles)    <- interpreter.invoke() + output read-out (yolact.rs:161-188), batched

Same argument meaning (640x480 frame of r<<24|g<<16|b<<8 pixels, mutated in place) and the same error
behaviour (the reference panics on every failure; here every failure raises TodError).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import Detections, TodError, YolactOptions, check, lib

DEFAULT_MODEL = "data/FRC_model.tflite"  # the CPU model of BASELINE config 1 (the reference hard-codes the edgetpu one, yolact.rs:19)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Yolact:
    def __init__(self, model_path=DEFAULT_MODEL, device=0, **options):
        o = YolactOptions()
        lib().tod_yolact_default_options(C.byref(o))
        for k, v in options.items():
            if not hasattr(o, k):
                raise TypeError("unknown option %r" % k)
            setattr(o, k, v)
        self.options = o
        h = C.c_void_p()
        check(lib().tod_yolact_create(str(model_path).encode(), int(device), C.byref(o), C.byref(h)))
        self._h = h
        self.device = device
        self.num_outputs = lib().tod_yolact_num_outputs(h)
        self.num_tensors = lib().tod_yolact_num_tensors(h)
        self.num_ops = lib().tod_yolact_num_ops(h)
        self.outputs = [self.output_info(i) for i in range(self.num_outputs)]

    @classmethod
    def init(cls, model_path=DEFAULT_MODEL, device=0, **options):
        """`Yolact::init()` (yolact.rs:17)."""
        return cls(model_path, device, **options)

    def close(self):
        if getattr(self, "_h", None):
            try:
                lib().tod_yolact_destroy(self._h)
            except Exception:  # interpreter shutdown: the module globals may already be gone
                pass
            self._h = None

    __del__ = close

    # ------------------------------------------------------------------ introspection
    def output_info(self, i):
        shape = (C.c_int32 * 4)()
        sc, zp, n = C.c_float(), C.c_int32(), C.c_int32()
        check(lib().tod_yolact_output_info(self._h, i, shape, C.byref(sc), C.byref(zp), C.byref(n)))
        return dict(shape=tuple(shape), scale=sc.value, zero_point=zp.value, elems=n.value)

    def tensor_info(self, t):
        shape = (C.c_int32 * 4)()
        ty, sc, zp, n = C.c_int32(), C.c_float(), C.c_int32(), C.c_int32()
        check(lib().tod_yolact_tensor_info(self._h, t, shape, C.byref(ty), C.byref(sc), C.byref(zp), C.byref(n)))
        return dict(shape=tuple(shape), type=ty.value, scale=sc.value, zero_point=zp.value, elems=n.value)

    def stats(self):
        m, l, t = C.c_int64(), C.c_int32(), C.c_int32()
        check(lib().tod_yolact_stats(self._h, C.byref(m), C.byref(l), C.byref(t)))
        return dict(macs_per_tile=m.value, launches_per_call=l.value, tc_conv_layers=t.value)

    def set_priors(self, priors):
        p = np.ascontiguousarray(priors, np.float32)
        check(lib().tod_yolact_set_priors(self._h, _ptr(p), p.shape[0]))

    # ------------------------------------------------------------------ the reference call
    def classify(self, frame_buffer, width=640, height=480):
        """In place on a u32[height*width] frame (or u32[n, height*width] batch).  Returns True when the
        reference itself would not have returned (literal flood fill)."""
        fb = frame_buffer
        if not (isinstance(fb, np.ndarray) and fb.dtype == np.uint32 and fb.flags.c_contiguous and fb.flags.writeable):
            raise TypeError("classify needs a writable, contiguous uint32 array (it is mutated in place)")
        if fb.size % (width * height) != 0 or fb.size == 0:
            # the reference panics in copy_from_slice on a length mismatch (yolact.rs:233)
            raise TodError(-1, "frame buffer holds %d pixels, not a multiple of %dx%d" % (fb.size, width, height))
        n = fb.size // (width * height)
        rc = check(lib().tod_yolact_classify_batch(self._h, _ptr(fb), n, width, height))
        return rc == _lib.TOD_WARN_REFERENCE_DIVERGES

    def classify_device(self, d_frames_ptr, n, width, height, d_target_ptr=None, stream=None):
        check(lib().tod_yolact_classify_batch_device(self._h, d_frames_ptr, n, width, height, d_target_ptr, stream))

    # ------------------------------------------------------------------ batched tile inference
    def infer_tiles(self, tiles, outputs=True, tile_classes=True, detections=False, masks=True, float_masks=True, tile_masks=False,
                    cell_classes=False):
        """tiles: u8[n, th, tw, 3].  Returns dict(outputs=[u8 arrays], tile_classes=u32[n,th,tw], dets=..., diverged=bool);
        cell_classes=True adds the 28x28 grid the tile map is an 8x replication of (1/64 of the read-back)."""
        tiles = np.ascontiguousarray(tiles, np.uint8)
        n, th, tw = tiles.shape[0], tiles.shape[1], tiles.shape[2]
        res = {}
        outs = None
        if outputs:
            outs = [np.zeros((n,) + o["shape"][1:], np.uint8) for o in self.outputs]
            arr = (C.c_void_p * self.num_outputs)(*[a.ctypes.data for a in outs])
        tc = np.zeros((n, th, tw), np.uint32) if tile_classes else None
        det, keep = (None, None)
        if detections:
            det, keep = self._alloc_dets(n, masks, tile_hw=(th, tw) if tile_masks else None)
            if not float_masks:  # binary masks only: the library then skips the sigmoid (sign of the integer logit)
                det.masks = None
                keep["masks"] = None
        cells = None
        if cell_classes:
            gh, gw = self.outputs[4]["shape"][1], self.outputs[4]["shape"][2]
            cells = np.zeros((n, gh, gw), np.uint32)
        rc = check(lib().tod_yolact_infer_tiles_cells(self._h, _ptr(tiles), n, arr if outputs else None, _ptr(tc), _ptr(cells),
                                                      C.byref(det) if det is not None else None))
        res["outputs"] = outs
        res["tile_classes"] = tc
        res["cell_classes"] = cells
        res["diverged"] = rc == _lib.TOD_WARN_REFERENCE_DIVERGES
        if detections:
            res["dets"] = self._unpack_dets(keep, n)
        return res

    def infer_tiles_device(self, d_tiles_ptr, n, stream=None):
        check(lib().tod_yolact_infer_tiles_device(self._h, d_tiles_ptr, n, stream))

    def fetch_output(self, index, n):
        out = np.zeros((n,) + self.outputs[index]["shape"][1:], np.uint8)
        check(lib().tod_yolact_fetch_output(self._h, index, n, _ptr(out)))
        return out

    def fetch_output_f32(self, index, n):
        """The reference's dequantised `results[index]` (yolact.rs:169-188), computed on the GPU."""
        out = np.zeros((n,) + self.outputs[index]["shape"][1:], np.float32)
        check(lib().tod_yolact_fetch_output_f32(self._h, index, n, _ptr(out)))
        return out

    def fetch_tensor(self, t, n):
        info = self.tensor_info(t)
        dt = {0: np.float32, 2: np.int32, 3: np.uint8, 9: np.int8}[info["type"]]
        out = np.zeros((n,) + info["shape"][1:], dt)
        check(lib().tod_yolact_fetch_tensor(self._h, t, n, _ptr(out), out.nbytes))
        return out

    def fetch_tile_classes(self, n, th=224, tw=224):
        out = np.zeros((n, th, tw), np.uint32)
        check(lib().tod_yolact_fetch_tile_classes(self._h, n, _ptr(out)))
        return out

    def fetch_detections(self, n, masks=True):
        det, keep = self._alloc_dets(n, masks)
        check(lib().tod_yolact_fetch_detections(self._h, n, C.byref(det)))
        return self._unpack_dets(keep, n)

    def last_diverged(self):
        d = C.c_int(0)
        check(lib().tod_yolact_last_diverged(self._h, C.byref(d)))
        return bool(d.value)

    def profile_ops(self, n):
        cap = 512
        ms = np.zeros(cap, np.float32)
        kinds = np.zeros(cap, np.int32)
        k = check(lib().tod_yolact_profile_ops(self._h, n, _ptr(ms), _ptr(kinds), cap))
        return ms[:k], kinds[:k]

    def trace_steps(self, n):
        cap = 600
        ms = np.zeros(cap, np.float32)
        lanes = np.zeros(cap, np.int32)
        kinds = np.zeros(cap, np.int32)
        k = check(lib().tod_yolact_trace_steps(self._h, n, _ptr(ms), _ptr(lanes), _ptr(kinds), cap))
        return ms[:k], lanes[:k], kinds[:k]

    def step_macs(self):
        cap = 512
        m = np.zeros(cap, np.int64)
        k = check(lib().tod_yolact_step_macs(self._h, _ptr(m), cap))
        return m[:k]

    # ------------------------------------------------------------------ helpers
    def _proto_hw(self):
        sp = [o["shape"] for i, o in enumerate(self.outputs) if o["shape"][1] > 1 and i != 4]
        return (sp[0][1], sp[0][2]) if sp else (56, 56)

    def _alloc_dets(self, n, masks, tile_hw=None):
        md = self.options.max_dets
        ph, pw = self._proto_hw()
        keep = dict(count=np.zeros(n, np.int32), boxes=np.zeros((n, md, 4), np.float32), scores=np.zeros((n, md), np.float32),
                    classes=np.zeros((n, md), np.int32), priors=np.zeros((n, md), np.int32),
                    masks=np.zeros((n, md, ph, pw), np.float32) if masks else None,
                    masks_bin=np.zeros((n, md, ph, pw), np.uint8) if masks else None,
                    masks_bits=np.zeros((n, md, (ph * pw + 31) // 32), np.uint32) if masks else None,
                    masks_tile_bits=np.zeros((n, md, (tile_hw[0] * tile_hw[1] + 31) // 32), np.uint32) if (masks and tile_hw) else None)
        det = Detections(md, *[keep[k].ctypes.data if keep[k] is not None else None
                               for k in ("count", "boxes", "scores", "classes", "priors", "masks", "masks_bin", "masks_bits", "masks_tile_bits")])
        return det, keep

    @staticmethod
    def _unpack_dets(keep, n):
        out = []
        for t in range(n):
            c = int(keep["count"][t])
            out.append(dict(n=c, box=keep["boxes"][t, :c], score=keep["scores"][t, :c], cls=keep["classes"][t, :c],
                            prior=keep["priors"][t, :c],
                            masks=keep["masks"][t, :c] if keep["masks"] is not None else None,
                            masks_bin=keep["masks_bin"][t, :c] if keep["masks_bin"] is not None else None,
                            masks_bits=keep["masks_bits"][t, :c] if keep.get("masks_bits") is not None else None,
                            masks_tile_bits=keep["masks_tile_bits"][t, :c] if keep.get("masks_tile_bits") is not None else None))
        return out


class YolactPool:
    """`depth` handles that take alternate batches, each from its own host thread: the frame loop's double / triple
    buffering.  One batch's host<->device copies and latency-bound tail (small pyramid levels, Fast-NMS, mask assembly)
    overlap the next batch's backbone.  Every batch still goes through one handle exactly as `Yolact.infer_tiles` would run
    it (same results); the C-ABI call blocks and ctypes drops the GIL, so plain threads are enough."""

    def __init__(self, model_path=DEFAULT_MODEL, device=0, depth=3, **options):
        from concurrent.futures import ThreadPoolExecutor
        options.setdefault("batches_in_flight", max(1, depth))  # the handles share the GPU (tod_yolact_options::batches_in_flight)
        self.handles = [Yolact(model_path, device, **options) for _ in range(max(1, depth))]
        self._workers = [ThreadPoolExecutor(1) for _ in self.handles]  # one thread per handle: a handle runs one batch at a time
        self._next = 0

    def set_priors(self, priors):
        for h in self.handles:
            h.set_priors(priors)

    def submit(self, tiles, **kw):
        """Queues one batch; returns a future of `Yolact.infer_tiles(tiles, **kw)`."""
        k = self._next % len(self.handles)
        self._next += 1
        return self._workers[k].submit(self.handles[k].infer_tiles, tiles, **kw)

    def map(self, batches, **kw):
        return [f.result() for f in [self.submit(b, **kw) for b in batches]]

    def close(self):
        for w in self._workers:
            w.shutdown(wait=True)
        for h in self.handles:
            h.close()


class Pool:
    """The C-ABI frame sharder (`tod_pool_*`, include/tod.h): `depth` handles on each device, one host thread per handle
    inside the library; n tiles / frames are cut into contiguous per-GPU ranges and chunks of `max_tiles`, and results land
    in the caller's arrays at the frame's index.  devices=None -> every visible GPU."""

    def __init__(self, model_path=DEFAULT_MODEL, devices=None, depth=3, **options):
        o = YolactOptions()
        lib().tod_yolact_default_options(C.byref(o))
        for k, v in options.items():
            if not hasattr(o, k):
                raise TypeError("unknown option %r" % k)
            setattr(o, k, v)
        self.options = o
        dev = (C.c_int32 * len(devices))(*devices) if devices else None
        h = C.c_void_p()
        check(lib().tod_pool_create(str(model_path).encode(), dev, len(devices) if devices else 0, int(depth), C.byref(o), C.byref(h)))
        self._h = h
        self.num_devices = lib().tod_pool_num_devices(h)
        self.num_handles = lib().tod_pool_num_handles(h)

    def close(self):
        if getattr(self, "_h", None):
            try:
                lib().tod_pool_destroy(self._h)
            except Exception:
                pass
            self._h = None

    __del__ = close

    def infer_tiles(self, tiles, shapes, tile_classes=False, cell_classes=True, detections=True, masks_bits=True):
        """tiles u8[n,th,tw,3]; shapes = Yolact.outputs of a handle on the same model (output shapes).  Returns
        dict(outputs, tile_classes, cell_classes, dets=dict of [n, max_dets, ...] arrays)."""
        tiles = np.ascontiguousarray(tiles, np.uint8)
        n, th, tw = tiles.shape[:3]
        outs = [np.zeros((n,) + o["shape"][1:], np.uint8) for o in shapes]
        arr = (C.c_void_p * len(outs))(*[a.ctypes.data for a in outs])
        tc = np.zeros((n, th, tw), np.uint32) if tile_classes else None
        gh, gw = shapes[4]["shape"][1], shapes[4]["shape"][2]
        cells = np.zeros((n, gh, gw), np.uint32) if cell_classes else None
        det, keep = None, None
        if detections:
            md = self.options.max_dets
            sp = [o["shape"] for i, o in enumerate(shapes) if o["shape"][1] > 1 and i != 4]
            ph, pw = sp[0][1], sp[0][2]
            keep = dict(count=np.zeros(n, np.int32), boxes=np.zeros((n, md, 4), np.float32), scores=np.zeros((n, md), np.float32),
                        classes=np.zeros((n, md), np.int32), priors=np.zeros((n, md), np.int32),
                        masks_bits=np.zeros((n, md, (ph * pw + 31) // 32), np.uint32) if masks_bits else None)
            det = Detections(md, keep["count"].ctypes.data, keep["boxes"].ctypes.data, keep["scores"].ctypes.data, keep["classes"].ctypes.data,
                             keep["priors"].ctypes.data, None, None, keep["masks_bits"].ctypes.data if masks_bits else None, None)
        rc = check(lib().tod_pool_infer_tiles(self._h, _ptr(tiles), n, arr, _ptr(tc), _ptr(cells), C.byref(det) if det is not None else None))
        return dict(outputs=outs, tile_classes=tc, cell_classes=cells, dets=keep, diverged=rc == _lib.TOD_WARN_REFERENCE_DIVERGES)

    def classify(self, frames, width=640, height=480):
        """`Yolact::classify` for every frame of u32[n, height*width], in place, sharded over the pool."""
        fb = frames
        if not (isinstance(fb, np.ndarray) and fb.dtype == np.uint32 and fb.flags.c_contiguous and fb.flags.writeable):
            raise TypeError("classify needs a writable, contiguous uint32 array (it is mutated in place)")
        if fb.size % (width * height) != 0 or fb.size == 0:
            raise TodError(-1, "frame buffer holds %d pixels, not a multiple of %dx%d" % (fb.size, width, height))
        rc = check(lib().tod_pool_classify_batch(self._h, _ptr(fb), fb.size // (width * height), width, height))
        return rc == _lib.TOD_WARN_REFERENCE_DIVERGES

    def rgbd(self, frames, depth, scene_params, want=("map", "world", "conn0", "conn1", "balls")):
        """The fused RGB-D loop (scene.rs:84-97 -> scene.rs:147-331) for n frames: frames u32[n, H*W] classified in place,
        depth u16[n, H, W]; returns the scene outputs as SceneBuilder.append_batch does."""
        W, H = scene_params.width, scene_params.height
        n = frames.size // (W * H)
        depth = np.ascontiguousarray(depth, np.uint16).reshape(n, H, W)
        out = dict(map=np.zeros((n, H, W), np.uint32) if "map" in want else None,
                   world=np.zeros((n, H, W, 4), np.float32) if "world" in want else None,
                   conn0=np.zeros((n, H, W, 4), np.float32) if "conn0" in want else None,
                   conn1=np.zeros((n, H, W, 4), np.float32) if "conn1" in want else None,
                   balls=np.zeros((n, 100, 4), np.float32) if "balls" in want else None)
        check(lib().tod_pool_rgbd_batch(self._h, C.byref(scene_params), _ptr(frames), _ptr(depth), n, _ptr(out["map"]), _ptr(out["world"]),
                                        _ptr(out["conn0"]), _ptr(out["conn1"]), _ptr(out["balls"])))
        return out


def model_inspect(path):
    no, nt, m = C.c_int32(), C.c_int32(), C.c_int64()
    check(lib().tod_model_inspect(str(path).encode(), C.byref(no), C.byref(nt), C.byref(m)))
    return dict(num_ops=no.value, num_tensors=nt.value, macs=m.value)

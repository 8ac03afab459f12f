"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/_build/liboracle.so (the CPU restatement of the reference's hot path).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h")) or f == "Makefile"]
    stale = not os.path.exists(_LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


class SceneParams(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("max_depth_in", C.c_float), ("x_fov", C.c_float),
                ("y_fov", C.c_float), ("bot_avoidance_const", C.c_float), ("bot_norm_const", C.c_int),
                ("terrain_norm_const", C.c_int), ("bump_err", C.c_float), ("sample_shift", C.c_int),
                ("weights_mode", C.c_int)]


class DetectCfg(C.Structure):
    _fields_ = [("num_priors", C.c_int), ("num_classes", C.c_int), ("mask_dim", C.c_int), ("proto_h", C.c_int),
                ("proto_w", C.c_int), ("conf_thresh", C.c_float), ("nms_thresh", C.c_float), ("top_k", C.c_int),
                ("max_dets", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.tod_oracle_last_error.restype = C.c_char_p
        L.tod_oracle_model_load.restype = C.c_void_p
        L.tod_oracle_model_load.argtypes = [C.c_char_p]
        L.tod_oracle_model_free.argtypes = [C.c_void_p]
        for f in ("num_tensors", "num_ops", "num_inputs", "num_outputs"):
            getattr(L, "tod_oracle_model_" + f).argtypes = [C.c_void_p]
        for f in ("output_tensor", "input_tensor", "op_code"):
            getattr(L, "tod_oracle_model_" + f).argtypes = [C.c_void_p, C.c_int]
        L.tod_oracle_model_op_output.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.tod_oracle_model_tensor_info.restype = C.c_int64
        L.tod_oracle_model_tensor_info.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.tod_oracle_model_invoke.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.tod_oracle_model_tensor_data.restype = C.c_void_p
        L.tod_oracle_model_tensor_data.argtypes = [C.c_void_p, C.c_int]
        L.tod_oracle_model_macs.restype = C.c_int64
        L.tod_oracle_model_macs.argtypes = [C.c_void_p]
        L.tod_oracle_srdhm.argtypes = [C.c_int32, C.c_int32]
        L.tod_oracle_rdivpot.argtypes = [C.c_int32, C.c_int]
        L.tod_oracle_mbqm.argtypes = [C.c_int32, C.c_int32, C.c_int]
        L.tod_oracle_bump_table.argtypes = [C.c_float, C.c_int, C.c_float, C.c_void_p]
        L.tod_oracle_dequant_u8.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p]
        L.tod_oracle_detect.argtypes = [C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_float, C.c_int, C.c_void_p, C.c_float, C.c_int,
                                        C.c_void_p, C.c_float, C.c_int, C.c_void_p, C.c_float, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


# ------------------------------------------------------------------ scene
def scene_params(**kw):
    p = SceneParams()
    lib().tod_oracle_scene_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def bump_table(val, s, bump_err=0.1):
    out = np.zeros((2 * s, 2 * s), np.uint32)
    lib().tod_oracle_bump_table(C.c_float(val), s, C.c_float(bump_err), _p(out))
    return out


def pt_cloud(depth, target, params=None):
    """depth, target: u16[H,W] -> (map u32[H,W], balls f32[100,4])"""
    p = params or scene_params()
    depth = _c(depth, np.uint16)
    target = _c(target, np.uint16)
    assert depth.shape == (p.height, p.width) and target.shape == depth.shape
    m = np.zeros((p.height, p.width), np.uint32)
    balls = np.zeros((100, 4), np.float32)
    lib().tod_oracle_pt_cloud(C.byref(p), _p(depth), _p(target), _p(m), _p(balls))
    return m, balls


def pt_cloud_weights(m, params=None):
    p = params or scene_params()
    m = _c(m, np.uint32)
    world = np.zeros((p.height, p.width, 4), np.float32)
    c0 = np.zeros_like(world)
    c1 = np.zeros_like(world)
    lib().tod_oracle_pt_cloud_weights(C.byref(p), _p(m), _p(world), _p(c0), _p(c1))
    return world, c0, c1


def path_modify(height, pos, balls, connections, width=640, height_px=480, cap=4096):
    """path.rs:25-120 in intent mode (see oracle/path.cpp): returns dict(cost, pred, directions [n,2] or None if unreachable)"""
    n = width * height_px
    h = _c(height, np.float32).reshape(n)
    p3 = _c(pos, np.float32).reshape(n, 3)
    b2 = _c(balls, np.int32).reshape(-1, 2)
    c8 = _c(connections, np.float32).reshape(n, 8)
    cost = np.zeros(n, np.float32)
    pred = np.zeros(n, np.int32)
    dirs = np.zeros((cap, 2), np.float32)
    L = lib()
    L.tod_oracle_path_modify.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 7 + [C.c_int]
    k = L.tod_oracle_path_modify(width, height_px, _p(h), _p(p3), _p(b2), _p(c8), _p(cost), _p(pred), _p(dirs), cap)
    return dict(cost=cost, pred=pred, directions=None if k < 0 else dirs[:min(k, cap)].copy(), n=k)


def path_serialize(created_secs, directions):
    d = _c(directions, np.float32).reshape(-1, 2)
    out = np.zeros(8 + 8 * len(d), np.uint8)
    L = lib()
    L.tod_oracle_path_serialize.argtypes = [C.c_uint64, C.c_void_p, C.c_int, C.c_void_p]
    L.tod_oracle_path_serialize(int(created_secs), _p(d), len(d), _p(out))
    return out.tobytes()


def scene_materialize(m, world, c0, c1, balls):
    n = m.size
    height = np.zeros(n, np.float32)
    pos = np.zeros((n, 3), np.float32)
    b2 = np.zeros((100, 2), np.int32)
    conn = np.zeros((n, 8), np.float32)
    lib().tod_oracle_scene_materialize(n, _p(_c(m, np.uint32)), _p(_c(world, np.float32)), _p(_c(c0, np.float32)),
                                       _p(_c(c1, np.float32)), _p(_c(balls, np.float32)), _p(height), _p(pos), _p(b2),
                                       _p(conn))
    return height, pos, b2, conn


# ------------------------------------------------------------------ yolact.rs literal path
def dequant_u8(q, scale, zp):
    q = _c(q, np.uint8)
    out = np.zeros(q.shape, np.float32)
    lib().tod_oracle_dequant_u8(_p(q), q.size, C.c_float(scale), int(zp), _p(out))
    return out


def cell_classes(seg):
    seg = _c(seg, np.float32).reshape(-1, seg.shape[-1])
    out = np.zeros(seg.shape[0], np.uint8)
    lib().tod_oracle_cell_classes(_p(seg), seg.shape[0], seg.shape[1], _p(out))
    return out


def terrible_id(classes, mode=0):
    classes = _c(classes, np.uint8).reshape(784)
    ids = np.zeros(784, np.int8)
    diverges = lib().tod_oracle_terrible_id(_p(classes), mode, _p(ids))
    return ids, bool(diverges)


def pack_upsample(classes, ids, mode=0):
    out = np.zeros((224, 224), np.uint32)
    lib().tod_oracle_pack_upsample(_p(_c(classes, np.uint8)), _p(_c(ids, np.int8)), mode, _p(out))
    return out


def resize_triangle_rgb8(src, dw, dh):
    src = _c(src, np.uint8)
    sh, sw, _ = src.shape
    dst = np.zeros((dh, dw, 3), np.uint8)
    lib().tod_oracle_resize_triangle_rgb8(_p(src), sw, sh, _p(dst), dw, dh)
    return dst


def classify_pre(frame, width=640, height=480):
    frame = _c(frame, np.uint32).reshape(height * width)
    tiles = np.zeros((2, 224, 224, 3), np.uint8)
    lib().tod_oracle_classify_pre(_p(frame), width, height, _p(tiles))
    return tiles


def classify_post(t1, t2, width=640, height=480):
    out = np.zeros(height * width, np.uint32)
    lib().tod_oracle_classify_post(_p(_c(t1, np.uint32)), _p(_c(t2, np.uint32)), width, height, _p(out))
    return out


def target_from_frame(frame):
    frame = _c(frame, np.uint32)
    out = np.zeros(frame.shape, np.uint16)
    lib().tod_oracle_target_from_frame(_p(frame), frame.size, _p(out))
    return out


def postprocess_tile(seg_q, scale, zp, mode=0):
    """yolact.rs:90-131 on the quantised output #4 ([28,28,81] u8): returns (u32[224,224], diverges)"""
    seg = dequant_u8(seg_q, scale, zp).reshape(784, -1)
    cls = cell_classes(seg)
    ids, div = terrible_id(cls, mode)
    return pack_upsample(cls, ids, mode), div


# ------------------------------------------------------------------ detection
def detect_cfg(**kw):
    c = DetectCfg()
    lib().tod_oracle_detect_default_cfg(C.byref(c))
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def make_priors():
    out = np.zeros((3147, 4), np.float32)
    n = lib().tod_oracle_make_priors(_p(out), 3147)
    assert n == 3147
    return out


def detect(cls_q, cls_qp, box_q, box_qp, coef_q, coef_qp, proto_q, proto_qp, cfg=None, priors=None):
    """*_q: u8 arrays; *_qp: (scale, zero_point). Returns dict of detections."""
    cfg = cfg or detect_cfg()
    priors = make_priors() if priors is None else _c(priors, np.float32)
    md, ph, pw = cfg.max_dets, cfg.proto_h, cfg.proto_w
    box = np.zeros((md, 4), np.float32)
    score = np.zeros(md, np.float32)
    cls = np.zeros(md, np.int32)
    prior = np.zeros(md, np.int32)
    masks = np.zeros((md, ph, pw), np.float32)
    mbin = np.zeros((md, ph, pw), np.uint8)
    n = lib().tod_oracle_detect(C.byref(cfg), _p(priors),
                                _p(_c(cls_q, np.uint8)), C.c_float(cls_qp[0]), int(cls_qp[1]),
                                _p(_c(box_q, np.uint8)), C.c_float(box_qp[0]), int(box_qp[1]),
                                _p(_c(coef_q, np.uint8)), C.c_float(coef_qp[0]), int(coef_qp[1]),
                                _p(_c(proto_q, np.uint8)), C.c_float(proto_qp[0]), int(proto_qp[1]),
                                _p(box), _p(score), _p(cls), _p(prior), _p(masks), _p(mbin))
    return dict(n=n, box=box[:n], score=score[:n], cls=cls[:n], prior=prior[:n], masks=masks[:n], masks_bin=mbin[:n])


# ------------------------------------------------------------------ tflite graph
def upsample_masks(masks, out_h, out_w):
    """YOLACT `postprocess` (layers/output_utils.py, not in the reference; north-star row 9): the cropped prototype-resolution
    masks are resized to the tile with `F.interpolate(mode='bilinear', align_corners=False)` and thresholded at 0.5.
    float32 arithmetic in PyTorch's order: src = max(scale * (dst + 0.5) - 0.5, 0), i0 = int(src), i1 = i0 + (i0 < in - 1),
    l1 = src - i0, l0 = 1 - l1; v = h0 * (w0 * m[y0, x0] + w1 * m[y0, x1]) + h1 * (w0 * m[y1, x0] + w1 * m[y1, x1]).
    Returns (float32 [N, out_h, out_w], uint8 same shape)."""
    m = np.ascontiguousarray(masks, np.float32)
    n, ih, iw = m.shape
    f = np.float32

    def axis(out, inn):
        scale = f(inn) / f(out)
        src = np.maximum(scale * (np.arange(out, dtype=np.float32) + f(0.5)) - f(0.5), f(0))
        i0 = src.astype(np.int32)
        i1 = i0 + (i0 < inn - 1)
        l1 = (src - i0.astype(np.float32)).astype(np.float32)
        l0 = (f(1) - l1).astype(np.float32)
        return i0, i1, l0, l1

    y0, y1, h0, h1 = axis(out_h, ih)
    x0, x1, w0, w1 = axis(out_w, iw)
    top = (w0[None, None, :] * m[:, y0][:, :, x0] + w1[None, None, :] * m[:, y0][:, :, x1]).astype(np.float32)
    bot = (w0[None, None, :] * m[:, y1][:, :, x0] + w1[None, None, :] * m[:, y1][:, :, x1]).astype(np.float32)
    up = (h0[None, :, None] * top + h1[None, :, None] * bot).astype(np.float32)
    return up, (up > f(0.5)).astype(np.uint8)


class Model:
    def __init__(self, path):
        self._h = lib().tod_oracle_model_load(os.fsencode(path))
        if not self._h:
            raise RuntimeError("oracle: " + lib().tod_oracle_last_error().decode())
        L = lib()
        self.num_tensors = L.tod_oracle_model_num_tensors(self._h)
        self.num_ops = L.tod_oracle_model_num_ops(self._h)
        self.inputs = [L.tod_oracle_model_input_tensor(self._h, i) for i in range(L.tod_oracle_model_num_inputs(self._h))]
        self.outputs = [L.tod_oracle_model_output_tensor(self._h, i) for i in range(L.tod_oracle_model_num_outputs(self._h))]

    def __del__(self):
        if getattr(self, "_h", None):
            lib().tod_oracle_model_free(self._h)
            self._h = None

    def op_code(self, i):
        return lib().tod_oracle_model_op_code(self._h, i)

    def op_output(self, i, k=0):
        return lib().tod_oracle_model_op_output(self._h, i, k)

    def tensor_info(self, t):
        shape = (C.c_int * 4)()
        ty, sc, zp = C.c_int(), C.c_float(), C.c_int()
        n = lib().tod_oracle_model_tensor_info(self._h, t, shape, C.byref(ty), C.byref(sc), C.byref(zp))
        return dict(shape=tuple(shape), type=ty.value, scale=sc.value, zero_point=zp.value, elems=n)

    def invoke(self, tile_u8, threads=0):
        tile_u8 = _c(tile_u8, np.uint8)
        if lib().tod_oracle_model_invoke(self._h, _p(tile_u8), threads) != 0:
            raise RuntimeError("oracle: " + lib().tod_oracle_last_error().decode())

    def tensor(self, t):
        info = self.tensor_info(t)
        dt = {0: np.float32, 2: np.int32, 3: np.uint8, 4: np.int64, 9: np.int8}[info["type"]]
        ptr = lib().tod_oracle_model_tensor_data(self._h, t)
        n = info["elems"]
        buf = (C.c_char * (n * np.dtype(dt).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dt, count=n).reshape(info["shape"]).copy()

    def macs(self):
        return lib().tod_oracle_model_macs(self._h)

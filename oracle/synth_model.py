"""ORACLE — TEST INFRASTRUCTURE ONLY.

Deterministic generator of a synthetic stand-in for the missing `data/FRC_model.tflite`
(/root/reference/.MISSING_LARGE_BLOBS:1-2), plus an independent numpy restatement of the TFLite
integer kernels that is used (a) to calibrate activation ranges while generating and (b) to
cross-check the C++ oracle (tests/test_oracle_tflite.py).

Topology (inferred from /root/reference/data/FRC_model_edgetpu.log:7-19,
data/README.md:5,10 and src/yolact.rs:91,108,118): uint8 [1,224,224,3] -> MobileNetV2 (alpha 1,
truncated after block 14) -> FPN-256 (P3..P7) -> protonet (32 prototypes @56x56), shared
prediction head (3 anchors; 81 classes, 4 box, 32 coefficients + TANH) and a 1x1 semantic head
(28x28x81 == output #4).  Operator histogram equals the reference log exactly:
64 CONV_2D, 15 DEPTHWISE_CONV_2D, 19 QUANTIZE, 15 RESHAPE, 11 ADD, 5 TANH, 5 PAD,
3 RESIZE_BILINEAR, 3 CONCATENATION, 1 RELU = 141 ops.

Everything is reproducible bit-for-bit on any machine: weights come from a counter-based
splitmix64 stream, and activation ranges are calibrated with *integer-exact* arithmetic
(float64 BLAS matmuls over integer-valued matrices whose partial sums stay below 2^53, so the
summation order cannot matter).
"""
import ctypes
import hashlib
import math
import os
import struct
import sys

import numpy as np

from .fbwriter import Builder

# tflite schema enums
T_FLOAT32, T_INT32, T_UINT8, T_INT8 = 0, 2, 3, 9
OP_ADD, OP_CONCAT, OP_CONV, OP_DW, OP_RELU, OP_RESHAPE, OP_RESIZE, OP_TANH, OP_PAD, OP_QUANT = 0, 2, 3, 4, 19, 22, 23, 28, 34, 114
ACT_NONE, ACT_RELU, ACT_RELU6 = 0, 1, 3
PAD_SAME, PAD_VALID = 0, 1
# BuiltinOptions union tags
BO_CONV, BO_DW, BO_CONCAT, BO_ADD, BO_RESIZE, BO_RESHAPE, BO_PAD, BO_QUANT = 1, 2, 10, 11, 15, 17, 22, 85

_libm = ctypes.CDLL("libm.so.6")
_libm.tanhf.restype = ctypes.c_float
_libm.tanhf.argtypes = [ctypes.c_float]

f32 = np.float32


# ---------------------------------------------------------------- deterministic RNG
class Stream:
    """counter-based splitmix64; identical output on every machine"""

    def __init__(self, seed):
        self.seed = np.uint64(seed)
        self.ctr = 0

    def u64(self, n):
        with np.errstate(over="ignore"):
            i = np.arange(self.ctr, self.ctr + n, dtype=np.uint64)
            x = self.seed + (i + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
            x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            x = x ^ (x >> np.uint64(31))
        self.ctr += n
        return x

    def bytes4sum(self, n):
        """sum of four uniform bytes - 510: integer, roughly normal with sigma 147.8"""
        x = self.u64(n)
        s = (x & np.uint64(255)) + ((x >> np.uint64(8)) & np.uint64(255)) + ((x >> np.uint64(16)) & np.uint64(255)) + (
            (x >> np.uint64(24)) & np.uint64(255))
        return s.astype(np.int64) - 510

    def uniform(self, n):
        return (self.u64(n) >> np.uint64(11)).astype(np.float64) / float(1 << 53)


# ---------------------------------------------------------------- numpy restatement of TFLite integer rules
def quantize_multiplier(m):
    if m == 0.0:
        return 0, 0
    f, e = math.frexp(m)
    q = int(round_half_away(f * (1 << 31)))
    if q == (1 << 31):
        q //= 2
        e += 1
    if e < -31:
        e, q = 0, 0
    return q, e


def round_half_away(x):
    return math.floor(abs(x) + 0.5) * (1 if x >= 0 else -1)


def np_round_half_away(x):
    return np.where(x >= 0, np.floor(np.abs(x) + 0.5), -np.floor(np.abs(x) + 0.5))


def srdhm(a, b):
    a = a.astype(np.int64)
    ab = a * np.int64(b) if np.isscalar(b) else a * b.astype(np.int64)
    nudge = np.where(ab >= 0, np.int64(1 << 30), np.int64(1 - (1 << 30)))
    t = ab + nudge
    return np.where(t >= 0, t >> 31, -((-t) >> 31))  # truncating division by 2^31


def rdivpot(x, e):
    e = np.asarray(e, dtype=np.int64)
    mask = (np.int64(1) << e) - 1
    rem = x & mask
    thr = (mask >> 1) + (x < 0)
    return (x >> e) + (rem > thr)


def mbqm(x, q, shift):
    shift = np.asarray(shift, dtype=np.int64)
    left = np.maximum(shift, 0)
    right = np.maximum(-shift, 0)
    return rdivpot(srdhm(x.astype(np.int64) << left, q), right)


def act_range(act, scale, zp, ttype=T_INT8):
    qmin, qmax = (0, 255) if ttype == T_UINT8 else (-128, 127)
    quant = lambda v: zp + int(round_half_away(float(f32(v) / f32(scale))))
    lo, hi = qmin, qmax
    if act == ACT_RELU:
        lo = max(qmin, quant(0.0))
    elif act == ACT_RELU6:
        lo, hi = max(qmin, quant(0.0)), min(qmax, quant(6.0))
    return lo, hi


def conv_out_pad(padding, n, k, stride):
    if padding == PAD_SAME:
        out = (n + stride - 1) // stride
    else:
        out = (n + stride - k) // stride
    total = max(0, (out - 1) * stride + k - n)
    return out, (total // 2 if padding == PAD_SAME else 0)


def im2col(x, k, stride, pad, out_hw, fill):
    """x: [N,H,W,C] int64 (already offset by -zp); out-of-range taps contribute `fill` (0)."""
    N, H, W, C = x.shape
    OH, OW = out_hw
    ph, pw = pad
    xp = np.full((N, H + 2 * k, W + 2 * k, C), fill, dtype=x.dtype)
    xp[:, k:k + H, k:k + W] = x
    cols = []
    for fy in range(k):
        for fx in range(k):
            y0 = k - ph + fy
            x0 = k - pw + fx
            cols.append(xp[:, y0:y0 + stride * (OH - 1) + 1:stride, x0:x0 + stride * (OW - 1) + 1:stride])
    return cols  # list of [N,OH,OW,C]


def conv_acc(x, zp, w, stride, padding):
    """exact int accumulators. x [N,H,W,C] ints, w [OC,KH,KW,IC] ints -> int64 [N,OH,OW,OC]"""
    N, H, W, C = x.shape
    OC, KH, KW, IC = w.shape
    OH, ph = conv_out_pad(padding, H, KH, stride)
    OW, pw = conv_out_pad(padding, W, KW, stride)
    cols = im2col(x.astype(np.int64) - zp, KH, stride, (ph, pw), (OH, OW), 0)
    A = np.concatenate(cols, axis=-1).reshape(N * OH * OW, KH * KW * IC).astype(np.float64)
    B = w.reshape(OC, KH * KW * IC).astype(np.float64).T
    acc = A @ B  # exact: integer-valued, |partial sums| < 2^53
    return np.rint(acc).astype(np.int64).reshape(N, OH, OW, OC)


def dw_acc(x, zp, w, stride, padding):
    """w [1,KH,KW,C]"""
    N, H, W, C = x.shape
    _, KH, KW, _ = w.shape
    OH, ph = conv_out_pad(padding, H, KH, stride)
    OW, pw = conv_out_pad(padding, W, KW, stride)
    cols = im2col(x.astype(np.int64) - zp, KH, stride, (ph, pw), (OH, OW), 0)
    acc = np.zeros((N, OH, OW, C), np.int64)
    i = 0
    for fy in range(KH):
        for fx in range(KW):
            acc += cols[i] * w[0, fy, fx].astype(np.int64)
            i += 1
    return acc


def resize_bilinear_int(x, OH, OW, align_corners, half_pixel):
    N, IH, IW, C = x.shape

    def interp(n_out, n_in):
        s10 = ((1 << 10) * n_in + n_out // 2) // n_out
        if align_corners and n_out > 1:
            s10 = ((1 << 10) * (n_in - 1) + (n_out - 1) // 2) // (n_out - 1)
        v = np.arange(n_out, dtype=np.int64) * s10
        if half_pixel:
            v = v + (s10 // 2 - (1 << 9))
        trunc = lambda a: np.where(a >= 0, a // (1 << 10), -((-a) // (1 << 10)))  # C++ int division
        lo = np.maximum(trunc(v), 0)
        hi = np.minimum(trunc(v + (1 << 10) - 1), n_in - 1)
        return v, lo, hi

    iy, y0, y1 = interp(OH, IH)
    ix, x0, x1 = interp(OW, IW)
    wy1 = (iy - (1 << 10) * y0)[None, :, None, None]
    wy0 = (1 << 10) - wy1
    wx1 = (ix - (1 << 10) * x0)[None, None, :, None]
    wx0 = (1 << 10) - wx1
    X = x.astype(np.int64)
    g = lambda yy, xx: X[:, yy][:, :, xx]
    o20 = g(y0, x0) * wy0 * wx0 + g(y1, x0) * wy1 * wx0 + g(y0, x1) * wy0 * wx1 + g(y1, x1) * wy1 * wx1
    rnd = np.where(o20 > 0, 1 << 19, -(1 << 19))
    t = o20 + rnd
    return np.where(t >= 0, t // (1 << 20), -((-t) // (1 << 20)))


# ---------------------------------------------------------------- graph builder (writes the .tflite while executing on the calibration batch)
class TensorRec:
    def __init__(self, idx, shape, ttype, scale=None, zp=None, name="", buffer=0, qdim=0):
        self.idx, self.shape, self.type, self.scale, self.zp, self.name, self.buffer, self.qdim = idx, list(shape), ttype, scale, zp, name, buffer, qdim
        self.val = None  # calibration value (int64 ndarray) for activations


class Graph:
    def __init__(self, seed, calib):
        self.rng = Stream(seed)
        self.tensors = []
        self.buffers = [b""]  # buffer 0 = empty
        self.ops = []  # (code, inputs, outputs, opts dict)
        self.calib = calib
        self.shared = {}

    # -- plumbing
    def tensor(self, shape, ttype, scale=None, zp=None, name="", data=None, qdim=0):
        buf = 0
        if data is not None:
            self.buffers.append(data)
            buf = len(self.buffers) - 1
        t = TensorRec(len(self.tensors), shape, ttype, scale, zp, name, buf, qdim)
        self.tensors.append(t)
        return t

    def op(self, code, ins, outs, **opts):
        self.ops.append((code, [t.idx for t in ins], [t.idx for t in outs], opts))

    @staticmethod
    def qparams_from_range(lo, hi):
        """TFLite-style asymmetric int8 params from a real range (range always includes 0)."""
        lo, hi = min(lo, 0.0), max(hi, 0.0)
        if hi - lo < 1e-6:
            hi = lo + 1e-6
        scale = f32((hi - lo) / 255.0)
        zp = int(round_half_away(-128.0 - lo / float(scale)))
        return float(scale), max(-128, min(127, zp))

    # -- ops
    def input_u8(self, shape, scale, zp):
        t = self.tensor(shape, T_UINT8, scale, zp, "input")
        t.val = self.calib.astype(np.int64)
        return t

    def quantize(self, x, scale, zp, ttype, name):
        out = self.tensor(x.shape, ttype, scale, zp, name)
        q, sh = quantize_multiplier(float(f32(x.scale)) / float(f32(scale)))
        qmin, qmax = (0, 255) if ttype == T_UINT8 else (-128, 127)
        out.val = np.clip(mbqm(x.val - x.zp, q, sh) + zp, qmin, qmax)
        self.op(OP_QUANT, [x], [out])
        return out

    def pad(self, x, pads, name):
        (pt, pb), (pl, pr) = pads
        N, H, W, C = x.shape
        pd = self.tensor([4, 2], T_INT32, name=name + "/paddings", data=struct.pack("<8i", 0, 0, pt, pb, pl, pr, 0, 0))
        out = self.tensor([N, H + pt + pb, W + pl + pr, C], x.type, x.scale, x.zp, name)
        v = np.full((x.val.shape[0], H + pt + pb, W + pl + pr, C), x.zp, np.int64)
        v[:, pt:pt + H, pl:pl + W] = x.val
        out.val = v
        self.op(OP_PAD, [x, pd], [out])
        return out

    def _weights(self, key, shape, fan_in, gain, oc_axis):
        """int8 weights ~ N(0, 40) clipped, per-channel scale so that real sigma = gain/sqrt(fan_in)"""
        if key in self.shared:
            return self.shared[key]
        n = int(np.prod(shape))
        w = np.clip(np_round_half_away(self.rng.bytes4sum(n) * 0.27), -127, 127).astype(np.int64).reshape(shape)
        oc = shape[oc_axis]
        jitter = 1.0 + 0.5 * (self.rng.uniform(oc) - 0.5)
        wscale = (gain / math.sqrt(fan_in) / 40.0 * jitter).astype(np.float32)
        rec = (w, wscale)
        self.shared[key] = rec
        return rec

    def conv(self, x, oc, k, stride, padding, act, name, gain=1.0, depthwise=False, out_q=None, share=None, range_shrink=0.92):
        N, H, W, C = x.shape
        if depthwise:
            wshape, fan_in, oc_axis, oc = [1, k, k, C], k * k, 3, C
        else:
            wshape, fan_in, oc_axis = [oc, k, k, C], k * k * C, 0
        key = share or name
        w, wscale = self._weights(key + "/w", wshape, fan_in, gain, oc_axis)
        wkey = key + "/wtensor"
        if wkey in self.shared:
            wt = self.shared[wkey]
        else:
            wt = self.tensor(wshape, T_INT8, [float(s) for s in wscale], [0] * oc, key + "/weights", data=w.astype(np.int8).tobytes(), qdim=oc_axis)
            self.shared[wkey] = wt
        bscale = (f32(x.scale) * wscale).astype(np.float32)  # bias scale = in_scale * w_scale (float32 product)
        braw = self.rng.bytes4sum(oc) / 147.8 * 0.15  # real-valued bias ~ N(0, 0.15)
        bias = np_round_half_away(braw / bscale.astype(np.float64)).astype(np.int64)
        bt = self.tensor([oc], T_INT32, [float(s) for s in bscale], [0] * oc, name + "/bias", data=bias.astype(np.int32).tobytes())
        acc = (dw_acc if depthwise else conv_acc)(x.val, x.zp, w, stride, padding) + bias
        real = acc.astype(np.float64) * (float(f32(x.scale)) * wscale.astype(np.float64))
        if out_q is None:
            if act == ACT_RELU6:
                out_q = (float(f32(6.0 / 255.0)), -128)
            else:
                lo, hi = float(real.min()), float(real.max())
                if act == ACT_RELU:
                    lo = 0.0
                out_q = self.qparams_from_range(lo * range_shrink, hi * range_shrink)
        oscale, ozp = out_q
        out = self.tensor([N, acc.shape[1], acc.shape[2], oc], T_INT8, oscale, ozp, name)
        qs = [quantize_multiplier(float(f32(x.scale)) * float(ws) / float(f32(oscale))) for ws in wscale]
        q = np.array([a for a, _ in qs], np.int64)
        sh = np.array([b for _, b in qs], np.int64)
        lo, hi = act_range(act, oscale, ozp)
        out.val = np.clip(mbqm(acc, q, sh) + ozp, lo, hi)
        if depthwise:
            self.op(OP_DW, [x, wt, bt], [out], padding=padding, stride=stride, act=act, depth_mult=1)
        else:
            self.op(OP_CONV, [x, wt, bt], [out], padding=padding, stride=stride, act=act)
        return out

    def add(self, a, b, name):
        real = (a.val - a.zp) * float(f32(a.scale)) + (b.val - b.zp) * float(f32(b.scale))
        oscale, ozp = self.qparams_from_range(float(real.min()) * 0.95, float(real.max()) * 0.95)
        out = self.tensor(a.shape, T_INT8, oscale, ozp, name)
        twice_max = float(f32(2) * max(f32(a.scale), f32(b.scale)))
        m1, s1 = quantize_multiplier(float(f32(a.scale)) / twice_max)
        m2, s2 = quantize_multiplier(float(f32(b.scale)) / twice_max)
        mo, so = quantize_multiplier(twice_max / float(f32(1 << 20) * f32(oscale)))
        y1 = mbqm((a.val - a.zp) * (1 << 20), m1, s1)
        y2 = mbqm((b.val - b.zp) * (1 << 20), m2, s2)
        out.val = np.clip(mbqm(y1 + y2, mo, so) + ozp, -128, 127)
        self.op(OP_ADD, [a, b], [out], act=ACT_NONE)
        return out

    def resize2x(self, x, name):
        N, H, W, C = x.shape
        sz = self.tensor([2], T_INT32, name=name + "/size", data=struct.pack("<2i", 2 * H, 2 * W))
        out = self.tensor([N, 2 * H, 2 * W, C], T_INT8, x.scale, x.zp, name)
        out.val = resize_bilinear_int(x.val, 2 * H, 2 * W, False, True)
        self.op(OP_RESIZE, [x, sz], [out], align_corners=False, half_pixel=True)
        return out

    def relu(self, x, name):
        real_hi = float(((x.val - x.zp) * float(f32(x.scale))).max())
        oscale, ozp = self.qparams_from_range(0.0, real_hi * 0.95)
        out = self.tensor(x.shape, T_INT8, oscale, ozp, name)
        q, sh = quantize_multiplier(float(f32(x.scale)) / float(f32(oscale)))
        lo = max(-128, ozp + int(round_half_away(float(f32(0.0) / f32(oscale)))))
        out.val = np.clip(mbqm(x.val - x.zp, q, sh) + ozp, lo, 127)
        self.op(OP_RELU, [x], [out])
        return out

    def tanh(self, x, name):
        oscale, ozp = float(f32(1.0 / 128.0)), 0
        out = self.tensor(x.shape, T_INT8, oscale, ozp, name)
        inv = f32(1) / f32(oscale)
        lut = {}
        for v in range(-128, 128):
            deq = f32(x.scale) * f32(v - x.zp)
            tr = f32(_libm.tanhf(ctypes.c_float(float(deq))))
            resc = round_half_away(float(tr * inv))
            lut[v] = max(min(127, int(f32(resc) + f32(ozp))), -128)
        out.val = np.vectorize(lut.get)(x.val).astype(np.int64)
        self.op(OP_TANH, [x], [out])
        return out

    def reshape(self, x, shape, name):
        sh = self.tensor([len(shape)], T_INT32, name=name + "/shape", data=struct.pack("<%di" % len(shape), *shape))
        out = self.tensor(shape, x.type, x.scale, x.zp, name)
        out.val = x.val.reshape([x.val.shape[0]] + list(shape[1:]))
        self.op(OP_RESHAPE, [x, sh], [out], new_shape=shape)
        return out

    def concat(self, xs, axis, name):
        shape = list(xs[0].shape)
        shape[axis] = sum(t.shape[axis] for t in xs)
        out = self.tensor(shape, xs[0].type, xs[0].scale, xs[0].zp, name)
        out.val = np.concatenate([t.val for t in xs], axis=axis)
        self.op(OP_CONCAT, xs, [out], axis=axis, act=ACT_NONE)
        return out

    # -- serialisation
    def serialize(self, inputs, outputs, description):
        b = Builder()
        buf_offs = []
        for data in self.buffers:
            if data:
                dv = b.vector(data, len(data), 16)
                buf_offs.append(b.table([(0, "o", dv)]))
            else:
                buf_offs.append(b.table([]))
        codes = sorted({c for c, _, _, _ in self.ops})
        code_offs = [b.table([(0, "b", min(c, 127)), (2, "i", 1), (3, "i", c)]) for c in codes]
        t_offs = []
        for t in self.tensors:
            q = None
            if t.scale is not None:
                scales = t.scale if isinstance(t.scale, list) else [t.scale]
                zps = t.zp if isinstance(t.zp, list) else [t.zp]
                q = b.table([(2, "o", b.vector_of("f", [float(f32(s)) for s in scales])), (3, "o", b.vector_of("q", zps)), (6, "i", t.qdim)])
            t_offs.append(b.table([(0, "o", b.vector_of("i", t.shape)), (1, "b", t.type), (2, "I", t.buffer), (3, "o", b.string(t.name)), (4, "o", q)]))
        o_offs = []
        for code, ins, outs, o in self.ops:
            bo, bot = None, 0
            if code == OP_CONV:
                bot = BO_CONV
                bo = b.table([(0, "b", o["padding"]), (1, "i", o["stride"]), (2, "i", o["stride"]), (3, "b", o["act"]), (4, "i", 1), (5, "i", 1)])
            elif code == OP_DW:
                bot = BO_DW
                bo = b.table([(0, "b", o["padding"]), (1, "i", o["stride"]), (2, "i", o["stride"]), (3, "i", o["depth_mult"]), (4, "b", o["act"]), (5, "i", 1), (6, "i", 1)])
            elif code == OP_ADD:
                bot = BO_ADD
                bo = b.table([(0, "b", o["act"]), (1, "B", 1)])
            elif code == OP_CONCAT:
                bot = BO_CONCAT
                bo = b.table([(0, "i", o["axis"]), (1, "b", o["act"])])
            elif code == OP_RESIZE:
                bot = BO_RESIZE
                bo = b.table([(2, "B", int(o["align_corners"])), (3, "B", int(o["half_pixel"]))])
            elif code == OP_RESHAPE:
                bot = BO_RESHAPE
                bo = b.table([(0, "o", b.vector_of("i", o["new_shape"]))])
            elif code == OP_PAD:
                bot = BO_PAD
                bo = b.table([])
            elif code == OP_QUANT:
                bot = BO_QUANT
                bo = b.table([])
            o_offs.append(b.table([(0, "I", codes.index(code)), (1, "o", b.vector_of("i", ins)), (2, "o", b.vector_of("i", outs)), (3, "B", bot), (4, "o", bo)]))
        sg = b.table([(0, "o", b.vector_of_offsets(t_offs)), (1, "o", b.vector_of("i", [t.idx for t in inputs])),
                      (2, "o", b.vector_of("i", [t.idx for t in outputs])), (3, "o", b.vector_of_offsets(o_offs)), (4, "o", b.string("main"))])
        root = b.table([(0, "I", 3), (1, "o", b.vector_of_offsets(code_offs)), (2, "o", b.vector_of_offsets([sg])),
                        (3, "o", b.string(description)), (4, "o", b.vector_of_offsets(buf_offs))])
        return b.finish(root)


# ---------------------------------------------------------------- calibration images + topology
def calib_images(n=2, seed=0xCA11B):
    """smooth colour blobs + noise, uint8 [n,224,224,3]"""
    rng = Stream(seed)
    yy, xx = np.mgrid[0:224, 0:224].astype(np.float64)
    imgs = []
    for _ in range(n):
        img = np.zeros((224, 224, 3))
        for _b in range(12):
            u = rng.uniform(7)
            cx, cy, r = u[0] * 224, u[1] * 224, 8 + u[2] * 60
            col = u[3:6] * 255
            m = np.exp(-(((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * r * r)))
            img += m[..., None] * col * (0.4 + 0.6 * u[6])
        img += (rng.uniform(224 * 224 * 3).reshape(224, 224, 3) - 0.5) * 40
        imgs.append(np.clip(np.floor(img), 0, 255))
    return np.stack(imgs).astype(np.uint8)


def build_frc_model(seed=0xF2C, small=False):
    """small=True builds a 64x64-input, thin variant with the same op types for fast tests."""
    calib = calib_images()
    S = 64 if small else 224
    if small:
        calib = calib[:, :S, :S]
    g = Graph(seed, calib)
    wd = (lambda c: max(8, c // 4)) if small else (lambda c: c)
    fpn_c = 64 if small else 256
    ncls, nanch, mdim = 81, 3, 32

    x = g.input_u8([1, S, S, 3], float(f32(1.0 / 128.0)), 128)
    x = g.quantize(x, float(f32(1.0 / 128.0)), 0, T_INT8, "input_int8")
    # --- MobileNetV2 backbone (Keras layout: explicit PAD before every stride-2 3x3)
    x = g.pad(x, ((0, 1), (0, 1)), "Conv1_pad")
    x = g.conv(x, wd(32), 3, 2, PAD_VALID, ACT_RELU6, "Conv1", gain=1.6)
    x = g.conv(x, 0, 3, 1, PAD_SAME, ACT_RELU6, "expanded_conv_depthwise", gain=1.8, depthwise=True)
    x = g.conv(x, wd(16), 1, 1, PAD_SAME, ACT_NONE, "expanded_conv_project", gain=1.0)
    cfgs = [(24, 2, 2), (32, 3, 2), (64, 4, 2), (96, 3, 1), (160, 2, 2)]
    taps = {}
    bi = 0
    for c, n, s in cfgs:
        for i in range(n):
            bi += 1
            stride = s if i == 0 else 1
            inp = x
            name = "block_%d" % bi
            h = g.conv(x, inp.shape[3] * 6, 1, 1, PAD_SAME, ACT_RELU6, name + "_expand", gain=1.6)
            if stride == 2:
                h = g.pad(h, ((0, 1), (0, 1)), name + "_pad")
                h = g.conv(h, 0, 3, 2, PAD_VALID, ACT_RELU6, name + "_depthwise", gain=1.8, depthwise=True)
            else:
                h = g.conv(h, 0, 3, 1, PAD_SAME, ACT_RELU6, name + "_depthwise", gain=1.8, depthwise=True)
            h = g.conv(h, wd(c), 1, 1, PAD_SAME, ACT_NONE, name + "_project", gain=0.9)
            x = g.add(inp, h, name + "_add") if (stride == 1 and inp.shape[3] == h.shape[3]) else h
        taps[c] = x
    c3, c4, c5 = taps[32], taps[96], taps[160]
    # --- FPN
    l5 = g.conv(c5, fpn_c, 1, 1, PAD_SAME, ACT_NONE, "fpn_lat5")
    l4 = g.conv(c4, fpn_c, 1, 1, PAD_SAME, ACT_NONE, "fpn_lat4")
    l3 = g.conv(c3, fpn_c, 1, 1, PAD_SAME, ACT_NONE, "fpn_lat3")
    t4 = g.add(l4, g.resize2x(l5, "fpn_up5"), "fpn_add4")
    t3 = g.add(l3, g.resize2x(t4, "fpn_up4"), "fpn_add3")
    p5 = g.conv(l5, fpn_c, 3, 1, PAD_SAME, ACT_RELU, "fpn_pred5", gain=1.4)
    p4 = g.conv(t4, fpn_c, 3, 1, PAD_SAME, ACT_RELU, "fpn_pred4", gain=1.4)
    p3 = g.conv(t3, fpn_c, 3, 1, PAD_SAME, ACT_RELU, "fpn_pred3", gain=1.4)
    p6 = g.conv(p5, fpn_c, 3, 2, PAD_SAME, ACT_RELU, "fpn_down6", gain=1.4)
    p7 = g.conv(p6, fpn_c, 3, 2, PAD_SAME, ACT_RELU, "fpn_down7", gain=1.4)
    # --- protonet on P3
    h = p3
    for i in range(3):
        h = g.conv(h, fpn_c, 3, 1, PAD_SAME, ACT_RELU, "proto_conv%d" % i, gain=1.4)
    h = g.resize2x(h, "proto_up")
    h = g.conv(h, fpn_c, 3, 1, PAD_SAME, ACT_RELU, "proto_conv3", gain=1.4)
    h = g.conv(h, mdim, 1, 1, PAD_SAME, ACT_NONE, "proto_out", gain=1.0)
    proto = g.relu(h, "proto_relu")
    # --- shared prediction head over P3..P7
    cls_l, box_l, coef_l = [], [], []
    for li, p in enumerate([p3, p4, p5, p6, p7]):
        n = "head%d" % li
        hw = p.shape[1] * p.shape[2]
        t = g.conv(p, fpn_c, 3, 1, PAD_SAME, ACT_RELU, n + "_tower", gain=1.4, share="head_tower")
        c = g.conv(t, nanch * ncls, 3, 1, PAD_SAME, ACT_NONE, n + "_cls", gain=1.2, share="head_cls")
        b = g.conv(t, nanch * 4, 3, 1, PAD_SAME, ACT_NONE, n + "_box", gain=1.2, share="head_box")
        m = g.conv(t, nanch * mdim, 3, 1, PAD_SAME, ACT_NONE, n + "_coef", gain=0.6, share="head_coef")
        m = g.tanh(m, n + "_coef_tanh")
        cls_l.append(g.reshape(c, [1, hw * nanch, ncls], n + "_cls_reshape"))
        box_l.append(g.reshape(b, [1, hw * nanch, 4], n + "_box_reshape"))
        coef_l.append(g.reshape(m, [1, hw * nanch, mdim], n + "_coef_reshape"))

    def common(ts):
        lo = min(float(((t.val - t.zp) * float(f32(t.scale))).min()) for t in ts)
        hi = max(float(((t.val - t.zp) * float(f32(t.scale))).max()) for t in ts)
        return Graph.qparams_from_range(lo, hi)

    cs, cz = common(cls_l)
    cls_l = [g.quantize(t, cs, cz, T_INT8, t.name + "_rq") for t in cls_l]
    bs, bz = common(box_l)
    box_l = [g.quantize(t, bs, bz, T_INT8, t.name + "_rq") for t in box_l]
    # TANH outputs already share (1/128, 0); the last three levels still carry an (identity)
    # requantise so that the QUANTIZE count matches the reference's operator log (19).
    coef_l = coef_l[:2] + [g.quantize(t, t.scale, t.zp, T_INT8, t.name + "_rq") for t in coef_l[2:]]
    cls = g.concat(cls_l, 1, "cls_concat")
    box = g.concat(box_l, 1, "box_concat")
    coef = g.concat(coef_l, 1, "coef_concat")
    seg = g.conv(p3, ncls, 1, 1, PAD_SAME, ACT_NONE, "seg_head", gain=1.5)
    u8 = lambda t, n: g.quantize(t, t.scale, t.zp + 128, T_UINT8, n)
    outs = [u8(box, "out_box"), u8(cls, "out_cls"), u8(coef, "out_coef"), u8(proto, "out_proto"), u8(seg, "out_seg")]
    blob = g.serialize([g.tensors[0]], outs, "synthetic FRC YOLACT (tod-b200 oracle/synth_model.py, seed %#x)" % seed)
    return g, blob


DEFAULT_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "FRC_model_synth.tflite")
SMALL_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "FRC_model_synth_small.tflite")


def ensure_models(force=False):
    """Write both synthetic models under oracle/_build/ if missing. Returns (full_path, small_path)."""
    os.makedirs(os.path.dirname(DEFAULT_PATH), exist_ok=True)
    for path, small in ((SMALL_PATH, True), (DEFAULT_PATH, False)):
        stale = os.path.exists(path) and os.path.getmtime(path) < os.path.getmtime(os.path.abspath(__file__))
        if force or stale or not os.path.exists(path):
            _, blob = build_frc_model(small=small)
            tmp = path + ".tmp%d" % os.getpid()
            with open(tmp, "wb") as f:
                f.write(blob)
            os.replace(tmp, path)
    return DEFAULT_PATH, SMALL_PATH


if __name__ == "__main__":
    small = "--small" in sys.argv
    g, blob = build_frc_model(small=small)
    path = SMALL_PATH if small else DEFAULT_PATH
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "wb") as f:
        f.write(blob)
    from collections import Counter
    print(path, len(blob), "bytes sha256", hashlib.sha256(blob).hexdigest()[:16])
    print(Counter(c for c, _, _, _ in g.ops), len(g.ops), "ops")

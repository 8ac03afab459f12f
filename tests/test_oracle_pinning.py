"""Pins the CPU oracle with a third, independent restatement (every bit-exact claim rested on the repo's own
C++ oracle and its numpy twin).  torch is in the image, so:

 * int8 graph ops: every CONV_2D / DEPTHWISE_CONV_2D / ADD / QUANTIZE / RELU / PAD output the oracle produces on random
   micro graphs is recomputed from the oracle's own input tensors with exact integer accumulation in torch
   (float64 `conv2d` over integers < 2^53 is exact) and the literal gemmlowp arithmetic of SURVEY §10
   (SaturatingRoundingDoublingHighMul, RoundingDivideByPOT, QuantizeMultiplier), written here from the published rules,
   not imported from oracle/.  RESIZE_BILINEAR (version-sensitive integer variant) is held within one
   code of torch's float bilinear; TANH within one code of float tanh.
 * decode / Fast-NMS / mask assembly: a float64 torch transcription of upstream YOLACT (layers/box_utils.py::decode,
   jaccard; layers/functions/detection.py::detect + fast_nms; layers/output_utils.py::crop) on the oracle model's own
   head outputs; the oracle's table-driven float32 result must agree to 1e-5 and select the same detections.

Still unpinned against a real TFLite interpreter (tools/pin_with_tflite.py closes that wherever tflite_runtime exists).
"""
import math

import numpy as np
import pytest

import oracle
from oracle import synth_model as sm
from tests import synth

torch = pytest.importorskip("torch")
F = torch.nn.functional


# ----------------------------------------------------------------------------- literal fixed-point rules
def quantize_multiplier(m):
    if m == 0.0:
        return 0, 0
    f, e = math.frexp(m)
    q = int(math.floor(f * (1 << 31) + 0.5))        # TfLiteRound of a positive value
    if q == 1 << 31:
        q //= 2
        e += 1
    if e < -31:
        return 0, 0
    return q, e


def srdhm(a, b):
    """SaturatingRoundingDoublingHighMul on int64 arrays holding int32 values"""
    a = np.asarray(a, np.int64)
    b = np.asarray(b, np.int64)
    sat = (a == -(1 << 31)) & (b == -(1 << 31))
    ab = a * b
    nudge = np.where(ab >= 0, 1 << 30, 1 - (1 << 30)).astype(np.int64)
    t = ab + nudge
    out = np.where(t >= 0, t // (1 << 31), -((-t) // (1 << 31)))   # C++ division truncates towards zero
    return np.where(sat, (1 << 31) - 1, out)


def rdivpot(x, e):
    x = np.asarray(x, np.int64)
    e = np.asarray(e, np.int64)
    mask = (np.int64(1) << e) - 1
    rem = x & mask
    thr = (mask >> 1) + (x < 0)
    return (x >> e) + (rem > thr)


def mbqm(x, q, shift):
    shift = np.asarray(shift, np.int64)
    return rdivpot(srdhm(np.asarray(x, np.int64) * (np.int64(1) << np.maximum(shift, 0)), q), np.maximum(-shift, 0))


def act_limits(act, scale, zp):
    f32 = np.float32
    quant = lambda v: zp + int(math.floor(abs(float(f32(v) / f32(scale))) + 0.5) * (1 if v >= 0 else -1))
    lo, hi = -128, 127
    if act == sm.ACT_RELU:
        lo = max(lo, quant(0.0))
    elif act == sm.ACT_RELU6:
        lo, hi = max(lo, quant(0.0)), min(hi, quant(6.0))
    return lo, hi


def same_pad(n, k, stride):
    out = (n + stride - 1) // stride
    total = max(0, (out - 1) * stride + k - n)
    return out, total // 2, total - total // 2


# ----------------------------------------------------------------------------- graph
def micro_graph(S, c0, expand, dw_stride, oc, seed):
    """stem -> expand -> depthwise (s1 / padded s2) -> project -> residual ADD (stride 1) -> TANH branch, RELU branch,
    RESIZE 2x -> 3x3 conv -> QUANTIZE to uint8: every op type of the FRC graph except RESHAPE / CONCATENATION (aliases)."""
    calib = sm.calib_images()[:, :S, :S]
    g = sm.Graph(seed, calib)
    x = g.input_u8([1, S, S, 3], float(sm.f32(1.0 / 128.0)), 128)
    x = g.quantize(x, float(sm.f32(1.0 / 128.0)), 0, sm.T_INT8, "input_int8")
    x = g.pad(x, ((0, 1), (0, 1)), "stem_pad")
    x = g.conv(x, c0, 3, 2, sm.PAD_VALID, sm.ACT_RELU6, "stem", gain=1.6)
    b0 = g.conv(x, c0, 1, 1, sm.PAD_SAME, sm.ACT_NONE, "bneck_in", gain=1.0)
    h = g.conv(b0, c0 * expand, 1, 1, sm.PAD_SAME, sm.ACT_RELU6, "expand", gain=1.6)
    if dw_stride == 2:
        h = g.pad(h, ((0, 1), (0, 1)), "dw_pad")
        h = g.conv(h, 0, 3, 2, sm.PAD_VALID, sm.ACT_RELU6, "dw", gain=1.8, depthwise=True)
        h = g.conv(h, oc, 1, 1, sm.PAD_SAME, sm.ACT_NONE, "project", gain=0.9)
    else:
        h = g.conv(h, 0, 3, 1, sm.PAD_SAME, sm.ACT_RELU6, "dw", gain=1.8, depthwise=True)
        h = g.conv(h, c0, 1, 1, sm.PAD_SAME, sm.ACT_NONE, "project", gain=0.9)
        h = g.add(b0, h, "residual")
    t = g.tanh(g.conv(h, 8, 3, 1, sm.PAD_SAME, sm.ACT_NONE, "coef", gain=1.2), "coef_tanh")
    r = g.relu(g.conv(h, 8, 1, 1, sm.PAD_SAME, sm.ACT_NONE, "proto", gain=1.2), "proto_relu")
    u = g.resize2x(h, "up")
    u = g.conv(u, oc, 3, 1, sm.PAD_SAME, sm.ACT_RELU, "smooth", gain=1.4)
    outs = [g.quantize(v, v.scale, v.zp + 128, sm.T_UINT8, "out%d" % i) for i, v in enumerate((u, t, r))]
    return g, g.serialize([g.tensors[0]], outs, "tod-b200 oracle-pinning micro graph")


def buffer_of(g, t, dtype):
    return np.frombuffer(g.buffers[t.buffer], dtype)


def check_op(g, m, code, ins, outs, opts):
    T = g.tensors
    out_t = T[outs[0]]
    got = m.tensor(out_t.idx).astype(np.int64)
    x_t = T[ins[0]]
    x = m.tensor(x_t.idx).astype(np.int64)
    if code in (sm.OP_CONV, sm.OP_DW):
        w_t, b_t = T[ins[1]], T[ins[2]]
        dw = code == sm.OP_DW
        w = buffer_of(g, w_t, np.int8).astype(np.float64).reshape(w_t.shape)
        bias = buffer_of(g, b_t, np.int32).astype(np.int64)
        k, stride = w_t.shape[1], opts["stride"]
        xi = torch.from_numpy((x - x_t.zp).astype(np.float64)).permute(0, 3, 1, 2)   # (in + input_offset), NCHW
        if opts["padding"] == sm.PAD_SAME:
            _, pt, pb = same_pad(x.shape[1], k, stride)
            _, pl, pr = same_pad(x.shape[2], k, stride)
            xi = F.pad(xi, (pl, pr, pt, pb))   # a zero after the offset == a skipped tap
        if dw:
            wt = torch.from_numpy(w).permute(3, 0, 1, 2)          # [1,k,k,C] -> [C,1,k,k]
            acc = F.conv2d(xi, wt, stride=stride, groups=w_t.shape[3])
        else:
            wt = torch.from_numpy(w).permute(0, 3, 1, 2)          # OHWI -> OIHW
            acc = F.conv2d(xi, wt, stride=stride)
        acc = acc.permute(0, 2, 3, 1).numpy()
        assert np.array_equal(acc, np.rint(acc)) and np.abs(acc).max() < 2 ** 31
        acc = acc.astype(np.int64) + bias
        qs = [quantize_multiplier(float(np.float32(x_t.scale)) * float(np.float32(ws)) / float(np.float32(out_t.scale))) for ws in w_t.scale]
        q = np.array([a for a, _ in qs], np.int64)
        sh = np.array([b for _, b in qs], np.int64)
        lo, hi = act_limits(opts["act"], out_t.scale, out_t.zp)
        want = np.clip(mbqm(acc, q, sh) + out_t.zp, lo, hi)
        return np.array_equal(got, want), int((got != want).sum())
    if code == sm.OP_ADD:
        y_t = T[ins[1]]
        y = m.tensor(y_t.idx).astype(np.int64)
        f32 = np.float32
        twice_max = float(f32(2) * max(f32(x_t.scale), f32(y_t.scale)))
        m1 = quantize_multiplier(float(f32(x_t.scale)) / twice_max)
        m2 = quantize_multiplier(float(f32(y_t.scale)) / twice_max)
        mo = quantize_multiplier(twice_max / float(f32(1 << 20) * f32(out_t.scale)))
        s1 = mbqm((x - x_t.zp) * (1 << 20), *m1)
        s2 = mbqm((y - y_t.zp) * (1 << 20), *m2)
        want = np.clip(mbqm(s1 + s2, *mo) + out_t.zp, -128, 127)
        return np.array_equal(got, want), int((got != want).sum())
    if code in (sm.OP_QUANT, sm.OP_RELU):
        q, sh = quantize_multiplier(float(np.float32(x_t.scale)) / float(np.float32(out_t.scale)))
        lo, hi = (0, 255) if out_t.type == sm.T_UINT8 else (-128, 127)
        if code == sm.OP_RELU:
            lo = max(lo, out_t.zp)
        want = np.clip(mbqm(x - x_t.zp, q, sh) + out_t.zp, lo, hi)
        return np.array_equal(got, want), int((got != want).sum())
    if code == sm.OP_PAD:
        pd = buffer_of(g, T[ins[1]], np.int32)
        want = np.pad(x, ((0, 0), (pd[2], pd[3]), (pd[4], pd[5]), (0, 0)), constant_values=x_t.zp)
        return np.array_equal(got, want), int((got != want).sum())
    if code == sm.OP_RESIZE:
        xf = torch.from_numpy(x.astype(np.float64)).permute(0, 3, 1, 2)
        want = F.interpolate(xf, size=tuple(out_t.shape[1:3]), mode="bilinear", align_corners=False).permute(0, 2, 3, 1).numpy()
        bad = int((np.abs(got - want) > 1.0).sum())
        return bad == 0, bad
    if code == sm.OP_TANH:
        real = np.tanh((x - x_t.zp).astype(np.float64) * float(np.float32(x_t.scale)))
        want = np.clip(np.rint(real * 128.0), -128, 127)
        bad = int((np.abs(got - want) > 1).sum())
        return bad == 0, bad
    return None, 0


@pytest.mark.parametrize("S,c0,expand,dw_stride,oc,seed", [(24, 16, 6, 1, 24, 11), (30, 8, 3, 2, 16, 12), (16, 32, 1, 1, 48, 13), (40, 16, 6, 2, 32, 14)])
def test_graph_ops_against_exact_integer_arithmetic(tmp_path, S, c0, expand, dw_stride, oc, seed):
    g, blob = micro_graph(S, c0, expand, dw_stride, oc, seed)
    path = tmp_path / "g.tflite"
    path.write_bytes(blob)
    m = oracle.Model(str(path))
    checked = {}
    for tile in synth.rgb_tiles(2, S=S, seed=seed):
        m.invoke(tile, threads=2)
        for code, ins, outs, opts in g.ops:
            ok, bad = check_op(g, m, code, ins, outs, opts)
            if ok is None:
                continue
            assert ok, "op code %d -> tensor '%s': %d elements differ from the exact restatement" % (code, g.tensors[outs[0]].name, bad)
            checked[code] = checked.get(code, 0) + 1
    want_codes = {sm.OP_CONV, sm.OP_DW, sm.OP_QUANT, sm.OP_PAD, sm.OP_RESIZE, sm.OP_TANH, sm.OP_RELU}
    if dw_stride == 1:
        want_codes.add(sm.OP_ADD)
    assert want_codes <= set(checked)


def test_full_model_convolutions_against_exact_integer_arithmetic(models):
    """The same recomputation on the layers bench.py times: every CONV_2D, DEPTHWISE and ADD of the full FRC stand-in
    (64 + 15 + 11 ops) on one tile, from the parsed file's own weights and quantisation parameters."""
    full, _ = models
    g, blob = sm.build_frc_model()
    assert blob == open(full, "rb").read(), "the model file on disk is not what synth_model generates"
    m = oracle.Model(full)
    m.invoke(synth.rgb_tiles(1, seed=21)[0], threads=8)
    n = {}
    for code, ins, outs, opts in g.ops:
        if code not in (sm.OP_CONV, sm.OP_DW, sm.OP_ADD):
            continue
        ok, bad = check_op(g, m, code, ins, outs, opts)
        assert ok, "full model op code %d -> '%s': %d elements differ" % (code, g.tensors[outs[0]].name, bad)
        n[code] = n.get(code, 0) + 1
    assert n == {sm.OP_CONV: 64, sm.OP_DW: 15, sm.OP_ADD: 11}


# ----------------------------------------------------------------------------- detection head
def upstream_detect(cls, box, coef, proto, priors, conf_thresh=0.05, nms_thresh=0.5, top_k=200, max_dets=100):
    """float64 transcription of dbolya/yolact: Detect.detect + fast_nms(second_threshold) + postprocess mask assembly + crop"""
    conf = torch.softmax(cls, -1)                                       # yolact.py eval: F.softmax(conf, -1)
    var0, var1 = 0.1, 0.2                                               # box_utils.decode
    b = torch.cat((priors[:, :2] + box[:, :2] * var0 * priors[:, 2:], priors[:, 2:] * torch.exp(box[:, 2:] * var1)), 1)
    b[:, :2] -= b[:, 2:] / 2
    b[:, 2:] += b[:, :2]
    cur = conf[:, 1:].t()                                               # [classes, priors]
    keep = cur.max(0).values > conf_thresh
    idx_all = torch.nonzero(keep)[:, 0]
    scores, boxes = cur[:, keep], b[keep]
    s_sorted, order = torch.sort(scores, dim=1, descending=True, stable=True)
    order, s_sorted = order[:, :top_k], s_sorted[:, :top_k]
    bb = boxes[order]                                                   # [C, k, 4]
    lt = torch.max(bb[:, :, None, :2], bb[:, None, :, :2])
    rb = torch.min(bb[:, :, None, 2:], bb[:, None, :, 2:])
    inter = (rb - lt).clamp(min=0).prod(-1)
    area = ((bb[..., 2] - bb[..., 0]) * (bb[..., 3] - bb[..., 1]))
    union = area[:, :, None] + area[:, None, :] - inter
    iou = torch.where(union > 0, inter / union, torch.zeros_like(inter))
    iou = iou.triu(diagonal=1)
    iou_max = iou.max(1).values
    k = (iou_max <= nms_thresh) & (s_sorted > conf_thresh)
    classes = torch.arange(cur.shape[0])[:, None].expand_as(k)[k]
    pri = idx_all[order][k]
    sc = s_sorted[k]
    sc, o2 = torch.sort(sc, dim=0, descending=True, stable=True)
    o2 = o2[:max_dets]
    sc, classes, pri = sc[:max_dets], classes[o2], pri[o2]
    out_boxes = b[pri]
    masks = torch.sigmoid(proto @ coef[pri].t())                        # [ph, pw, n]
    ph, pw = proto.shape[:2]
    x1, x2 = out_boxes[:, 0] * pw, out_boxes[:, 2] * pw                 # output_utils.crop, sanitize_coordinates(padding=1, cast=False)
    y1, y2 = out_boxes[:, 1] * ph, out_boxes[:, 3] * ph
    xa, xb = (torch.min(x1, x2) - 1).clamp(min=0), (torch.max(x1, x2) + 1).clamp(max=pw)
    ya, yb = (torch.min(y1, y2) - 1).clamp(min=0), (torch.max(y1, y2) + 1).clamp(max=ph)
    cols = torch.arange(pw, dtype=torch.float64)[None, :, None]
    rows = torch.arange(ph, dtype=torch.float64)[:, None, None]
    crop = (cols >= xa) & (cols < xb) & (rows >= ya) & (rows < yb)
    masks = masks * crop
    return dict(score=sc.numpy(), cls=classes.numpy(), prior=pri.numpy(), box=out_boxes.numpy(), masks=masks.permute(2, 0, 1).numpy())


@pytest.mark.parametrize("seed", [44, 45])
def test_detection_against_upstream_yolact_transcription(models, seed):
    full, _ = models
    m = oracle.Model(full)
    tile = synth.rgb_tiles(1, seed=seed)[0] if seed != 45 else sm.calib_images()[0]
    m.invoke(tile, threads=8)
    outs = [m.tensor(o) for o in m.outputs]
    info = [m.tensor_info(o) for o in m.outputs]
    qp = lambda i: (info[i]["scale"], info[i]["zero_point"])
    got = oracle.detect(outs[1], qp(1), outs[0], qp(0), outs[2], qp(2), outs[3], qp(3))
    deq = lambda i: torch.from_numpy((outs[i].astype(np.float64) - info[i]["zero_point"]) * float(np.float32(info[i]["scale"])))
    pri = torch.from_numpy(oracle.make_priors().astype(np.float64))
    want = upstream_detect(deq(1).reshape(-1, 81), deq(0).reshape(-1, 4), deq(2).reshape(-1, 32), deq(3).reshape(56, 56, 32), pri)
    n = got["n"]
    assert n == len(want["score"]) and n > 0
    np.testing.assert_allclose(np.sort(got["score"]), np.sort(want["score"]), rtol=2e-5)
    g_pairs = {(int(c), int(p)) for c, p in zip(got["cls"], got["prior"])}
    w_pairs = {(int(c), int(p)) for c, p in zip(want["cls"], want["prior"])}
    # float32 tables vs float64 softmax may swap detections only among scores tied with the last kept one
    cut = float(want["score"].min())
    tied = int((np.abs(want["score"] - cut) <= 2e-5 * cut).sum()) + int((np.abs(got["score"] - cut) <= 2e-5 * cut).sum())
    assert len(g_pairs ^ w_pairs) <= tied, "Fast-NMS selected different detections: %s" % sorted(g_pairs ^ w_pairs)[:6]
    w_index = {(int(c), int(p)): i for i, (c, p) in enumerate(zip(want["cls"], want["prior"]))}
    common = 0
    for i in range(n):
        key = (int(got["cls"][i]), int(got["prior"][i]))
        if key not in w_index:
            continue
        j = w_index[key]
        common += 1
        np.testing.assert_allclose(got["box"][i], want["box"][j], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(got["score"][i], want["score"][j], rtol=2e-5)
        np.testing.assert_allclose(got["masks"][i], want["masks"][j], rtol=1e-5, atol=1e-6)
    assert common >= n - tied

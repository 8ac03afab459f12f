"""Known answers for the detection oracle (YOLACT decode / Fast-NMS / masks; not in the reference: parity unpinned)."""
import pytest
import numpy as np

import oracle


def _head(P=3147, C=81):
    cls = np.zeros((P, C), np.uint8)
    cls[:, 0] = 255                       # background wins everywhere
    box = np.full((P, 4), 128, np.uint8)  # zero offsets (zp = 128)
    coef = np.full((P, 32), 128, np.uint8)
    proto = np.zeros((56, 56, 32), np.uint8)
    return cls, box, coef, proto


def test_priors_layout():
    p = oracle.make_priors()
    assert p.shape == (3147, 4)
    assert np.allclose(p[0], [0.5 / 28, 0.5 / 28, 12 / 224, 12 / 224])
    assert np.allclose(p[1, 2:], [12 * np.sqrt(0.5) / 224, 12 / np.sqrt(0.5) / 224])
    assert np.allclose(p[-1, :2], [0.75, 0.75]) and np.allclose(p[3 * 784, :2], [0.5 / 14, 0.5 / 14])


def test_fast_nms_suppression_and_order():
    cls, box, coef, proto = _head()
    qp = (0.1, 0)
    # three priors of the same cell (same centre, different aspect) + one far away, all class 5
    for p, v in ((0, 250), (1, 240), (2, 230), (3 * 400, 220)):
        cls[p, 0] = 100
        cls[p, 5] = v
    d = oracle.detect(cls, qp, box, (0.05, 128), coef, (1 / 128, 128), proto, (0.05, 0))
    # Fast-NMS: candidate j is dropped if ANY higher-scored candidate (even a suppressed one) overlaps > 0.5
    pri = oracle.make_priors()

    def iou(a, b):
        ax1, ay1, ax2, ay2 = a[0] - a[2] / 2, a[1] - a[3] / 2, a[0] + a[2] / 2, a[1] + a[3] / 2
        bx1, by1, bx2, by2 = b[0] - b[2] / 2, b[1] - b[3] / 2, b[0] + b[2] / 2, b[1] + b[3] / 2
        iw, ih = max(0, min(ax2, bx2) - max(ax1, bx1)), max(0, min(ay2, by2) - max(ay1, by1))
        return iw * ih / (a[2] * a[3] + b[2] * b[3] - iw * ih)
    keep = [0]
    for j in (1, 2):
        if max(iou(pri[i], pri[j]) for i in range(j)) <= 0.5:
            keep.append(j)
    keep.append(1200)
    assert list(d["prior"]) == keep and (d["cls"] == 4).all()
    assert (np.diff(d["score"]) <= 0).all()


def test_thresholds_and_masks():
    cls, box, coef, proto = _head()
    cls[10, 0] = 200
    cls[10, 7] = 255   # strong detection
    cls[20, 0] = 255
    cls[20, 9] = 200   # exp(-5.5) relative weight: below conf_thresh
    proto[..., 0] = 100
    coef[10, 0] = 255  # positive logit everywhere -> mask = box crop
    d = oracle.detect(cls, (0.1, 0), box, (0.05, 128), coef, (1 / 128, 128), proto, (0.05, 0))
    assert list(d["prior"]) == [10] and d["cls"][0] == 6
    m = d["masks_bin"][0]
    x1, y1, x2, y2 = d["box"][0] * 56
    ys, xs = np.nonzero(m)
    assert xs.min() == max(0, int(np.ceil(x1 - 1))) and ys.min() == max(0, int(np.ceil(y1 - 1)))
    assert m.sum() > 0 and (d["masks"][0][m == 1] > 0.5).all() and (d["masks"][0][m == 0] == 0).all()


def test_upsample_masks_is_torch_interpolate():
    """oracle.upsample_masks pins itself on the library YOLACT calls: F.interpolate(bilinear, align_corners=False) on CPU."""
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(11)
    m = rng.random((6, 56, 56)).astype(np.float32)
    m[:, :10] = 0.0  # cropped regions are exact zeros
    up, bits = oracle.upsample_masks(m, 224, 224)
    ref = torch.nn.functional.interpolate(torch.from_numpy(m)[None], (224, 224), mode="bilinear", align_corners=False)[0].numpy()
    np.testing.assert_allclose(up, ref, rtol=0, atol=2.5e-7)
    clear = np.abs(ref - 0.5) > 1e-6
    assert np.array_equal(bits[clear], (ref > 0.5)[clear].astype(np.uint8))
    up2, _ = oracle.upsample_masks(m[:, :7, :7], 7, 7)  # identity size: the source pixels come back exactly
    assert np.array_equal(up2, m[:, :7, :7])

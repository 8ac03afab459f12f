#!/usr/bin/env python
"""Runs one convolution shape through the tcgen05 kernel (and the CUDA-core cross-check): the short program
that `ncu --set full -k regex:conv_tc_kernel` profiles.   python tools/conv_probe.py tiles H W IC OC K [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tod_b200  # noqa: E402
from tod_b200 import _lib  # noqa: E402

a = [int(v) for v in sys.argv[1:]]
shape, iters = a[:6], (a[6] if len(a) > 6 else 3)
ms_tc, ms_direct, bad = _lib.conv_selftest(*shape, iters=iters)
t, h, w, ic, oc, k = shape
ops = 2.0 * t * h * w * ic * oc * k * k
print("shape %s: tcgen05 %.4f ms (%.1f TOP/s), cuda-core %.4f ms, mismatching bytes %d" % (shape, ms_tc, ops / ms_tc / 1e9, ms_direct, bad))

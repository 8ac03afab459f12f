#!/usr/bin/env python
"""Per-layer bounds of the tcgen05 convolutions of one step: geometry (TOD_TC_PLAN), measured time (back-to-back launches
between one event pair) and what each hardware limit would allow - tensor pipe, HBM bytes, TMA box rows, epilogue issue.

    python tools/layer_table.py [--tiles 64] > profiles/layers_rNN.txt
"""
import argparse
import os
import re
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["TOD_TC_PLAN"] = "1"
os.environ.setdefault("TOD_PROFILE_REPS", "10")
import numpy as np  # noqa: E402

import tod_b200  # noqa: E402
from oracle import synth_model  # noqa: E402
from tests import synth  # noqa: E402

PEAK_TOPS, HBM_GBPS, SMS, CLK = 4555.0, 6542.0, 148, 1.965e9
ROW_CYC, EPI_INSTR = 7.0, 7.8  # measured: TMA box rows per SM; epilogue SASS instructions per output byte


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiles", type=int, default=64)
    a = ap.parse_args()
    full, _ = synth_model.ensure_models()
    with tempfile.TemporaryFile() as tf:
        saved = os.dup(2)
        os.dup2(tf.fileno(), 2)
        try:
            y = tod_b200.Yolact.init(full, max_tiles=a.tiles)
        finally:
            os.dup2(saved, 2)
        tf.seek(0)
        plan = [dict((k, v) for k, v in re.findall(r"(\w+)=(\S+)", ln)) for ln in tf.read().decode().splitlines() if ln.startswith("TC_PLAN")]
    y.infer_tiles(synth.rgb_tiles(a.tiles), outputs=False, tile_classes=False)
    ms, kinds = y.profile_ops(a.tiles)
    ms2, _ = y.profile_ops(a.tiles)
    ms = np.minimum(ms, ms2)
    tc = [i for i, k in enumerate(kinds) if k & 0x1000]
    assert len(tc) == len(plan), (len(tc), len(plan))
    n = a.tiles
    print("# tiles=%d; us = measured (%s back-to-back launches / event pair); bounds in us: mma (%.0f TOP/s), hbm (%.0f GB/s), "
          "rows (%.0f cyc per TMA box row per SM), epi (%.1f instr per output byte, 4 issue/clk/SM); eff = max(bound) / us"
          % (n, os.environ["TOD_PROFILE_REPS"], PEAK_TOPS, HBM_GBPS, ROW_CYC, EPI_INSTR))
    print("%3s %-13s %-13s %1s %1s %-9s %4s %3s %5s %7s | %7s %6s %6s %6s %6s %5s" %
          ("op", "in", "out", "k", "s", "kernel", "BN", "BK", "mtile", "ctas", "us", "mma", "hbm", "rows", "epi", "eff"))
    tot = dict(us=0.0, mma=0.0, hbm=0.0, rows=0.0, epi=0.0, best=0.0)
    for i, pl in zip(tc, plan):
        ih, iw, ic = map(int, pl["in"].split("x"))
        oh, ow, oc = map(int, pl["out"].split("x"))
        k, s = int(pl["k"]), int(pl["s"])
        mt, nt, bn, bk, kch = int(pl["m_tiles"]), int(pl["n_tiles"]), int(pl["BN"]), int(pl["BK"]), int(pl["kchunks"])
        pw, ph, pn = map(int, pl["patch"].split("x"))
        macs = n * oh * ow * oc * ic * k * k
        byts = n * (ih * iw * ic + oh * ow * oc) + oc * ic * k * k
        mma = 2 * macs / PEAK_TOPS / 1e6
        hbm = byts / HBM_GBPS / 1e3
        ctas = min(mt * nt, SMS)
        waves = -(-mt * nt // ctas)
        # TMA box rows per work item: A rows per tap and K chunk (cp.async layers: none), B rows, output rows per store pass
        wo = int(pl["wo"])
        a_rows = 0 if pl["cp"] != "0" else (128 if pl["flat"] == "1" else ph * pn * (s if k == 3 else 1)) * k * k * kch
        if pl["flat"] == "0" and k == 3:
            a_rows = (ph * s) * pn * k * k * kch * (pw * s if s == 2 else 1)  # element-strided boxes fetch pixel by pixel
        b_rows = bn * k * k * kch / (2 if pl["pair"] == "1" else 1)
        o_rows = (128 * -(-min(bn, oc) // wo)) if wo else 0
        rows = waves * (a_rows + b_rows + o_rows) * ROW_CYC / CLK * 1e6
        epi = n * oh * ow * oc * EPI_INSTR / 32 / (4 * ctas) / CLK * 1e6
        kern = ("pair" if pl["pair"] == "1" else "fast" if pl["fast"] == "1" else "gen") + ("/cp" if pl["cp"] != "0" else "") + "/m%s" % pl["mode"]
        us = ms[i] * 1e3
        best = max(mma, hbm, rows, epi)
        print("%3d %-13s %-13s %1d %1d %-9s %4d %3d %5d %7d | %7.1f %6.1f %6.1f %6.1f %6.1f %5.2f" %
              (i, pl["in"], pl["out"], k, s, kern, bn, bk, mt, ctas, us, mma, hbm, rows, epi, max(mma, hbm) / us))
        for key, v in (("us", us), ("mma", mma), ("hbm", hbm), ("rows", rows), ("epi", epi), ("best", best)):
            tot[key] += v
    print("# totals: us %.1f  mma %.1f  hbm %.1f  rows %.1f  epi %.1f  sum(max bound) %.1f" % tuple(tot[k] for k in ("us", "mma", "hbm", "rows", "epi", "best")))
    other = [(i, float(ms[i]) * 1e3) for i, k in enumerate(kinds) if not (k & 0x1000)]
    print("# other steps (us): " + " ".join("%d:%.1f" % t for t in other))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Summarise an `ncu --set full` report of the conv_tc launches of one bench step: DRAM bytes and time per launch,
tensor-pipe activity.  Writes the JSON bench.py reads for roofline.traffic.

    python tools/ncu_traffic.py gpurun_out/conv_tc_step_r1_raw.csv profiles/conv_tc_traffic.json [profiles/conv_tc_step_r1.csv]
"""
import csv
import io
import json
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "%": 1.0}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    if rep.endswith(".csv"):  # already exported on the GPU box with `ncu -i x.ncu-rep --page raw --csv`
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    want = {"name": "Kernel Name", "grid": "Grid Size", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum", "t": "gpu__time_duration.sum",
            "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_el": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"}

    def val(r, key):
        i = col[want[key]]
        return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)

    launches = []
    for r in data:
        if "conv_tc" not in r[col["Kernel Name"]]:
            continue
        launches.append(dict(grid=r[col["Grid Size"]], dram_read=val(r, "rd"), dram_write=val(r, "wr"), us=val(r, "t"),
                             tensor_active_pct=val(r, "tensor"), tensor_elapsed_pct=val(r, "tensor_el")))
    n = len(launches)
    tot = sum(l["dram_read"] + l["dram_write"] for l in launches)
    summary = {"source": "ncu --set full --clock-control none, every conv_tc launch of one bench step (%s)" % rep.split("/")[-1], "launches": n,
               "dram_bytes_per_launch": tot / max(n, 1), "dram_read_bytes_per_step": sum(l["dram_read"] for l in launches),
               "dram_write_bytes_per_step": sum(l["dram_write"] for l in launches), "us_per_step_under_ncu": sum(l["us"] for l in launches),
               "tensor_pipe_active_pct_time_weighted": sum(l["tensor_elapsed_pct"] * l["us"] for l in launches) / max(1e-9, sum(l["us"] for l in launches)),
               "top_launch": max(launches, key=lambda l: l["us"]) if launches else None}
    json.dump(summary, open(out, "w"), indent=1)
    if len(sys.argv) > 3:
        with open(sys.argv[3], "w") as f:
            w = csv.writer(f)
            w.writerow(["launch", "grid", "us", "dram_read_bytes", "dram_write_bytes", "tensor_pipe_active_pct_of_elapsed"])
            for i, l in enumerate(launches):
                w.writerow([i, l["grid"], "%.2f" % l["us"], int(l["dram_read"]), int(l["dram_write"]), "%.2f" % l["tensor_elapsed_pct"]])
    print(json.dumps(summary))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Per-kernel summary of an `ncu --set full` raw CSV export (ncu -i x.ncu-rep --page raw --csv): time, DRAM bytes and
throughput, issue / pipe utilisation, occupancy, shared-memory wavefronts and the top stall reasons.

    python tools/ncu_summary.py gpurun_out/scene_r2_final_raw.csv profiles/scene_r2_ncu_summary.csv
"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "sm__cycles_elapsed.max"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    out = [["kernel"] + ["%s [%s]" % (k, units[col[k]]) for k in KEYS if k in col] + ["top stalls (warps per issue)"]]
    for d in data:
        name = d[col["Kernel Name"]].split("(")[0].split("::")[-1]
        st = []
        for h in stalls:
            try:
                st.append((float(d[col[h]].replace(",", "")), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
        st.sort(reverse=True)
        out.append([name] + [d[col[k]] for k in KEYS if k in col] + ["; ".join("%s %.2f" % (n, v) for v, n in st[:5] if n != "selected")])
    w = csv.writer(open(sys.argv[2], "w"))
    # transposed: one column per kernel reads better for a handful of kernels
    for r in zip(*out):
        w.writerow(r)
    for r in zip(*out):
        print(" | ".join(str(x)[:42] for x in r))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the handful of metrics DESIGN.md cites.
usage: python profiles/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
    "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum",
    "smsp__sass_inst_executed_op_global_ld.sum", "smsp__sass_inst_executed_op_global_st.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fp64_pred_on.sum",
    "sm__inst_executed_pipe_fp64.sum", "sm__cycles_elapsed.max",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("kernel:", d.get("Kernel Name"), " grid", d.get("Grid Size"), " block", d.get("Block Size"))
        for k in KEYS:
            if k in d:
                print("  %-70s %s %s" % (k, d[k], units[hdr.index(k)]))
        st = [(float(d[h]), h) for h in hdr
              if "issue_stalled" in h and h.endswith("per_warp_active.pct") and d[h]]
        for v, h in sorted(st, reverse=True)[:6]:
            print("  stall %-64s %.1f %%" % (h.replace("smsp__warps_issue_stalled_", "").replace("_per_warp_active.pct", ""), v))


if __name__ == "__main__":
    main(sys.argv[1])

#!/usr/bin/env python
"""Reduces an `ncu --set full` report to the numbers quoted in DESIGN.md / profiles/ (run here, no GPU needed):
   python tools/ncu_summary.py gpurun_out/r2_spatial.ncu-rep [pixels-per-launch] > profiles/r2_spatial_ncu_summary.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    ("time", "gpu__time_duration.sum"),
    ("dram read", "dram__bytes_read.sum"),
    ("dram write", "dram__bytes_write.sum"),
    ("warp instr", "smsp__inst_executed.sum"),
    ("thread instr", "smsp__thread_inst_executed.sum"),
    ("regs", "launch__registers_per_thread"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("static smem / block", "launch__shared_mem_per_block_static"),
    ("occupancy %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("issue slots busy %", "sm__inst_issued.avg.pct_of_peak_sustained_active"),
    ("pipe fma %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("pipe fmaheavy %", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active"),
    ("pipe alu %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("pipe xu %", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    ("pipe lsu %", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("pipe uniform %", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active"),
    ("L1 hit %", "l1tex__t_sector_hit_rate.pct"),
    ("L2 hit %", "lts__t_sector_hit_rate.pct"),
    ("dram throughput % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L1/TEX throughput %", "l1tex__throughput.avg.pct_of_peak_sustained_active"),
    ("L2 throughput %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L1 global load sectors", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"),
    ("L1 global load requests", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"),
    ("L2 sectors", "lts__t_sectors.sum"),
    ("SM throughput %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("smem loads (wavefronts)", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum"),
]
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
STALL_NAMES = ["long_scoreboard", "short_scoreboard", "wait", "not_selected", "selected", "no_instruction", "math_pipe_throttle", "mio_throttle", "lg_throttle", "barrier",
               "branch_resolving", "dispatch_stall", "drain", "imc_miss", "tex_throttle", "membar", "sleeping", "misc"]


def main():
    rep = sys.argv[1]
    pixels = float(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head = rows[0]
    units = rows[1]
    col = {n: i for i, n in enumerate(head)}
    print("ncu --set full --clock-control none, report %s (not committed); reduced with tools/ncu_summary.py" % rep)
    for r in rows[2:]:
        print("\n== %s" % r[col["Kernel Name"]][:150])
        for label, metric in WANT:
            if metric in col and r[col[metric]] != "":
                print("   %-28s %s %s" % (label, r[col[metric]], units[col[metric]]))
        if pixels and "smsp__thread_inst_executed.sum" in col:
            print("   %-28s %.0f" % ("thread instr / pixel", float(r[col["smsp__thread_inst_executed.sum"]].replace(",", "")) / pixels))
        st = []
        for n in STALL_NAMES:
            m = STALLS % n
            if m in col and r[col[m]] != "":
                st.append((float(r[col[m]].replace(",", "")), n))
        st.sort(reverse=True)
        print("   %-28s %s" % ("top stalls (warps/issue)", ", ".join("%s %.2f" % (n, v) for v, n in st[:6])))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of metrics we track.  Usage: ncu_summary.py file.ncu-rep [kernel-regex]"""
import csv, io, re, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_op_global_red.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
STALLS = "smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio"


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kcol = hdr.index("Kernel Name")
    for d in data:
        if len(sys.argv) > 2 and not re.search(sys.argv[2], d[kcol]):
            continue
        print("==", d[kcol][:80])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:72s} {d[i]:>18s} {units[i]}")
        st = []
        for i, h in enumerate(hdr):
            m = re.match(STALLS, h)
            if m:
                try:
                    st.append((float(d[i]), m.group(1)))
                except ValueError:
                    pass
        st.sort(reverse=True)
        print("  stalls (warps per issue):", ", ".join(f"{n}={v:.2f}" for v, n in st[:7]))


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Extract of an ncu report (raw page) for profiles/: python tools/ncu_extract.py report.ncu-rep "comment line" > profiles/xxx.csv
One line per (launch, metric) for the metrics the roofline discussion uses."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed.sum.per_cycle_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_lsu.sum", "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_xu.sum",
        "smsp__inst_executed_pipe_uniform.sum", "smsp__inst_executed_pipe_cbu.sum", "smsp__inst_executed_pipe_adu.sum"]
PREFIX = ["smsp__pcsamp_warps_issue_stalled_"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    if len(sys.argv) > 2:
        print("# " + sys.argv[2])
    print("launch,metric,value,unit")
    name_i = hdr.index("Kernel Name")
    for k, r in enumerate(rows[2:]):
        print(f'{k},Kernel Name,"{r[name_i]}",')
        for i, h in enumerate(hdr):
            if h in WANT or (any(h.startswith(p) for p in PREFIX) and "not_issued" not in h):
                print(f"{k},{h},{r[i]},{units[i]}")


if __name__ == "__main__":
    main()

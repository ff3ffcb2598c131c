#!/usr/bin/env python3
"""Summarise an `ncu --set full` report into the handful of counters DESIGN.md / bench.py cite.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-substring] > profiles/rN_….txt

Reads the report with `ncu -i … --page raw --csv` (no GPU needed) and prints, for every profiled
launch whose name contains the substring, the duration, DRAM traffic, pipe utilisation, issue
rate, occupancy and the top stall reasons."""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "sm__cycles_elapsed.avg.per_second",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes.sum.per_second",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__t_bytes.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, unit = rows[0], rows[1]
    col = {k: i for i, k in enumerate(head)}
    print(f"# source: {rep} (ncu --set full --clock-control none); summary by tools/ncu_summary.py")
    for row in rows[2:]:
        name = row[col["Kernel Name"]]
        if sub and sub not in name:
            continue
        if row[col.get("dram__bytes_read.sum", 0)] in ("-nan", "nan", ""):
            continue  # a launch captured without the full set
        print(f"\nKernel Name: {name}")
        print(f"  grid {row[col['Grid Size']]} block {row[col['Block Size']]} launch id {row[col['ID']]}")
        for k in KEYS:
            if k in col and row[col[k]] not in ("", "-nan"):
                print(f"  {k:78s} {row[col[k]]:>22s} {unit[col[k]]}")
        stalls = []
        for k, i in col.items():
            if k.startswith(STALL) and k.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(row[i]), k[len(STALL):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("  top stall reasons (warps stalled per issue-active cycle):")
        for v, k in stalls[:6]:
            print(f"    {k:40s} {v:8.3f}")


if __name__ == "__main__":
    main()

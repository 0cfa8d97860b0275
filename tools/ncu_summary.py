"""Print the handful of metrics we steer by from an ncu report:  python tools/ncu_summary.py rep.ncu-rep [more...]"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "_per_issue_active.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
        "launch__shared_mem_per_block_static", "launch__occupancy_limit", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed_pipe_fp64", "sm__inst_executed_pipe_fp64.sum"]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"## {rep}: {r[h.index('Kernel Name')][:70]}")
        for i, name in enumerate(h):
            if any(w in name for w in WANT):
                if name.startswith("dram__bytes") and ("pct" in name or "per_second" in name):
                    continue
                if r[i] in ("0", ""):
                    continue
                print(f"{name:95s} {units[i]:16s} {r[i]}")
        print()

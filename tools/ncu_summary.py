"""Key counters of every kernel in an ncu report: python tools/ncu_summary.py X.ncu-rep"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "memory_l1_wavefronts_shared", "memory_l1_wavefronts_shared_ideal",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "smsp__warps_eligible.avg.per_cycle_active",
] + [f"smsp__average_warps_issue_stalled_{k}_per_issue_active.ratio" for k in (
    "barrier", "short_scoreboard", "long_scoreboard", "math_pipe_throttle", "wait", "not_selected", "mio_throttle", "lg_throttle",
    "dispatch_stall", "branch_resolving", "no_instruction", "membar", "sleeping")]
names = [r[hdr.index("Kernel Name")] for r in rows[2:]]
print("kernels:", names)
for k in KEYS:
    if k in hdr:
        i = hdr.index(k)
        print(f"{k:85s} {rows[1][i]:8s}", [r[i] for r in rows[2:]])

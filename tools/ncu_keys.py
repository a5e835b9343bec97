"""Prints the metrics that matter from an `ncu --page raw --csv` export (one column block per kernel launch)."""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.avg.per_cycle_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max",
        "smsp__cycles_active.avg", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_not_selected_per_warp_active.pct", "smsp__warp_issue_stalled_membar_per_warp_active.pct",
        "smsp__warp_issue_stalled_sleeping_per_warp_active.pct", "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct",
        "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct", "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct",
        "smsp__warp_issue_stalled_tex_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_drain_per_warp_active.pct",
        "smsp__warp_issue_stalled_imc_miss_per_warp_active.pct", "smsp__warp_issue_stalled_selected_per_warp_active.pct",
        "smsp__warp_issue_stalled_misc_per_warp_active.pct"]
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr = i
        break
names = rows[hdr]
units = rows[hdr + 1]
extra = sys.argv[2:]
for r in rows[hdr + 2:]:
    d = dict(zip(names, r))
    print("==", d.get("Kernel Name", "?")[:90], "grid", d.get("Grid Size"), "block", d.get("Block Size"))
    for k in KEYS + extra:
        if k in d:
            print(f"   {k:82s} {d[k]:>16s} {units[names.index(k)]}")

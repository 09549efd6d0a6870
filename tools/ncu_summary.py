"""Markdown summary of ncu reports: python tools/ncu_summary.py report.ncu-rep [...] > profiles/xyz.md
Reads `ncu -i <rep> --page raw --csv` (one row per captured launch) and prints the metrics the DESIGN / bench quote."""
import csv, io, subprocess, sys

METRICS = [
    ("duration (ncu: caches flushed, kernel serialised)", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("registers / thread", "launch__registers_per_thread"),
    ("dynamic shared memory / CTA", "launch__shared_mem_per_block_dynamic"),
    ("cycles elapsed", "sm__cycles_elapsed.max"),
    ("SM cycles active (avg)", "sm__cycles_active.avg"),
    ("FP64 pipe, % of active cycles (instruction-issue based: a DMMA counts as one instruction, see note)", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
    ("FP64 pipe, % of elapsed cycles", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("issue slots busy, % of active", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("warp instructions", "smsp__inst_executed.sum"),
    ("warps resident per SM (avg)", "sm__warps_active.avg.per_cycle_active"),
    ("DRAM read", "dram__bytes_read.sum"),
    ("DRAM write", "dram__bytes_write.sum"),
    ("DRAM throughput % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("local loads (spills)", "smsp__sass_inst_executed_op_local_ld.sum"),
    ("local stores (spills)", "smsp__sass_inst_executed_op_local_st.sum"),
    ("shared-memory bank conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    ("stall long_scoreboard / issue", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("stall short_scoreboard / issue", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
    ("stall wait / issue", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    ("stall math_pipe_throttle / issue", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
    ("stall barrier / issue", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    ("stall no_instruction / issue", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"),
]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    print(f"## `{rep}`\n")
    for r in rows[2:]:
        print(f"### `{r[ik]}`\n\n| metric | value |\n|---|---|")
        for label, m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"| {label} (`{m}`) | {r[i]} {units[i]} |")
        print()
